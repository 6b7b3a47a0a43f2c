"""Deterministic synthetic inputs / weights shared by the golden generator, the tests, smoke() and bench.py.

TEST INFRASTRUCTURE.  Everything is regenerated from integer seeds with CPU generators, so the same tensors
are obtained in the authoring container and on the GPU box (same image, same torch build).
"""
import hashlib

import numpy as np
import torch

SEED = 88  # the reference's SEED, utils/defaults.yaml:1


def make_image(batch, seed=SEED):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, 256, 256, generator=g)


def init_state_dict(template_sd, seed=SEED):
    """Deterministic, well-conditioned random weights for every key of a reference-layout state_dict."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in template_sd:
        t = template_sd[k]
        if k.endswith('num_batches_tracked'):
            out[k] = torch.zeros_like(t)
        elif k.endswith('running_mean'):
            out[k] = torch.randn(t.shape, generator=g) * 0.1
        elif k.endswith('running_var'):
            out[k] = torch.rand(t.shape, generator=g) * 0.5 + 0.75
        elif k == 'decoder.dense_coor' or k == 'decoder.unsample_layer.weight' or '.mano_' in k:
            out[k] = t.clone().float()      # asset-derived tensors (graph / MANO tables) keep their values
        elif t.dim() == 1:
            if k.endswith('.weight'):      # norm scales
                out[k] = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
            else:                           # biases / norm shifts
                out[k] = 0.05 * torch.randn(t.shape, generator=g)
        else:
            fan_in = t[0].numel()
            if 'position_embeddings' in k:
                out[k] = 0.5 * torch.randn(t.shape, generator=g)
            else:
                out[k] = torch.randn(t.shape, generator=g) * (1.6 / fan_in) ** 0.5
    return out


def checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def make_labels(batch, seed=SEED):
    """Synthetic labels of SURVEY 8(d) for calc_loss_GCN."""
    g = torch.Generator().manual_seed(seed + 1)
    r = lambda *s: torch.randn(*s, generator=g)
    u = lambda *s: torch.rand(*s, generator=g)
    return {'v3d_l': r(batch, 778, 3) * 0.05, 'v3d_r': r(batch, 778, 3) * 0.05,
            'v2d_l': u(batch, 778, 2) * 256, 'v2d_r': u(batch, 778, 2) * 256, 'root_rel': r(batch, 3) * 0.05}


def make_loss_assets(assets, mano_l, mano_r):
    from . import model_ref
    out = {}
    for side, m in (('left', mano_l), ('right', mano_r)):
        jr = m['J_regressor']
        jr = np.asarray(jr.todense()) if hasattr(jr, 'todense') else np.asarray(jr)
        out[side] = {'J21': model_ref.make_joint_regressor(torch.from_numpy(jr.astype(np.float32))),
                     'faces': torch.from_numpy(np.asarray(m['f']).astype(np.int64)),
                     'perm': [int(v) for v in assets[side + '_graph']['graph_perm']]}
    return out


def make_mano_inputs(batch, ncomps=45, seed=SEED):
    g = torch.Generator().manual_seed(seed + 2)
    return {'axis': torch.rand(batch, 3, generator=g), 'pose_pca': torch.rand(batch, ncomps, generator=g),
            'pose_axis': torch.rand(batch, 45, generator=g) * 0.8, 'shape': torch.rand(batch, 10, generator=g),
            'trans': torch.rand(batch, 3, generator=g), 'scale': torch.rand(batch, generator=g) + 0.5}


def make_mano_loss_weights(batch, seed=SEED):
    """Fixed random cotangents for ManoLayer gradient checks: loss = <v, wv> + <j, wj>."""
    g = torch.Generator().manual_seed(seed + 3)
    return torch.randn(batch, 778, 3, generator=g), torch.randn(batch, 21, 3, generator=g)


def add_mano_assets(prepared, mano_left, mano_right):
    """MANO tables the 'newgraph' tail needs (oracle/model_ref.newgraph_tail): raw dicts with the left-hand shapedirs flip of
    decoder_lijun_mano.py:171-173 applied, and the 21-joint regressor of common/utils/mano.py:48-79 (tips 745/317/445/556/673)."""
    out = dict(prepared)
    dicts = {'left': dict(mano_left), 'right': dict(mano_right)}
    sl, sr = np.asarray(dicts['left']['shapedirs'], np.float32), np.asarray(dicts['right']['shapedirs'], np.float32)
    if np.abs(sl[:, 0, :] - sr[:, 0, :]).sum() < 1:
        sl = sl.copy(); sl[:, 0, :] *= -1
        dicts['left']['shapedirs'] = sl
    jr21 = {}
    for side, m in dicts.items():
        jr = m['J_regressor']
        jr = np.asarray(jr.todense()) if hasattr(jr, 'todense') else np.asarray(jr)
        tips = np.zeros((5, jr.shape[1]), np.float32)
        for i, v in enumerate((745, 317, 445, 556, 673)):
            tips[i, v] = 1.0
        jr = np.concatenate((jr.astype(np.float32), tips))[[0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]]
        jr21[side] = torch.from_numpy(jr)
    out['mano'], out['mano_jr21'] = dicts, jr21
    return out


def make_newgraph_cotangents(batch, seed=SEED):
    """Fixed random cotangents for the 'newgraph' outputs: loss = sum_k <out_k, w_k> (the reference's mano_loss_GCN is not importable here)."""
    g = torch.Generator().manual_seed(seed + 4)
    r = lambda *s: torch.randn(*s, generator=g)
    return {'verts3d_left': r(batch, 778, 3), 'verts3d_right': r(batch, 778, 3), 'verts2d_left': r(batch, 778, 2) * 0.01,
            'verts2d_right': r(batch, 778, 2) * 0.01, 'mano_pose_left': r(batch, 48), 'mano_pose_right': r(batch, 48),
            'mano_shape_left': r(batch, 10), 'mano_shape_right': r(batch, 10), 'joints3d_left': r(batch, 21, 3), 'joints3d_right': r(batch, 21, 3),
            'root_rel': r(batch, 3), 'v3c_left': r(batch, 252, 3), 'v3c_right': r(batch, 252, 3)}


def flat_newgraph(out):
    result, params, hlist, other = out
    d = {'root_rel': other['root_rel'], 'length': other['length'], 'v3d_left': result['v3d_left'], 'v3d_right': result['v3d_right']}
    for side in ('left', 'right'):
        d['verts3d_' + side] = result['verts3d'][side]; d['verts2d_' + side] = result['verts2d'][side]
        d['scale_' + side] = params['scale'][side]; d['trans2d_' + side] = params['trans2d'][side]
        d['scalelength_' + side] = params['scalelength_' + side]
        d['v3c_' + side] = hlist[0]['verts3d'][side]; d['v2c_' + side] = hlist[0]['verts2d'][side]
        for k in ('joints3d', 'mano_pose', 'mano_shape'):
            d[k + '_' + side] = other['verts3d_MANO_list'][side][k]
    return d


def newgraph_loss(flat, cot):
    return sum((flat[k] * w.to(flat[k].device, flat[k].dtype)).sum() for k, w in cot.items())


def make_mano_loss_case(batch=2, seed=SEED):
    """Seeded (prediction, label) tensors for the mano_loss_GCN golden: shaped like the 'newgraph' outputs (decoder_lijun_mano.py:286-305)."""
    g = torch.Generator().manual_seed(seed + 5)
    r = lambda *s: torch.randn(*s, generator=g)
    pred = {'verts3d_left': r(batch, 778, 3) * 0.05, 'verts3d_right': r(batch, 778, 3) * 0.05,
            'verts2d_left': torch.rand(batch, 778, 2, generator=g) * 256, 'verts2d_right': torch.rand(batch, 778, 2, generator=g) * 256,
            'mano_pose_left': r(batch, 48) * 0.4, 'mano_pose_right': r(batch, 48) * 0.4,
            'mano_shape_left': r(batch, 10), 'mano_shape_right': r(batch, 10), 'root_rel': r(batch, 3) * 0.05}
    lab = make_labels(batch, seed)
    lab.update({'lp_gt': r(batch, 48) * 0.3, 'rp_gt': r(batch, 48) * 0.3, 'ls_gt': r(batch, 10), 'rs_gt': r(batch, 10)})
    return pred, lab


def mano_loss_inputs(pred):
    """(result, paramsDict, handDictList, otherInfo) structure around the flat prediction dict"""
    result = {'verts3d': {s: pred['verts3d_' + s] for s in ('left', 'right')}, 'verts2d': {s: pred['verts2d_' + s] for s in ('left', 'right')}}
    other = {'root_rel': pred['root_rel'],
             'verts3d_MANO_list': {s: {'mano_pose': pred['mano_pose_' + s], 'mano_shape': pred['mano_shape_' + s]} for s in ('left', 'right')}}
    return result, {}, [], other


def make_eval_case(batch=6, seed=SEED):
    """Seeded (prediction, ground truth) meshes for the evaluation-metric golden (apps/eval_interhand.py loop body): [B,778,3] metres.
    Predictions are a similarity transform of the ground truth plus 4 mm noise (so the Procrustes numbers differ from the plain ones);
    the left ground truth touches the right one on about a third of the vertices (contact-deviation term); the last sample has the
    hands 1 m apart (no contact -> NaN)."""
    g = torch.Generator().manual_seed(seed + 9)
    r = lambda *s: torch.randn(*s, generator=g)
    gt_r = r(batch, 778, 3) * 0.04 + r(batch, 1, 3) * 0.1
    perm = torch.stack([torch.randperm(778, generator=g) for _ in range(batch)])
    gt_l = torch.gather(gt_r, 1, perm[:, :, None].repeat(1, 1, 3)) + r(batch, 778, 3) * 0.0012
    far = torch.rand(batch, 778, 1, generator=g) > 0.35
    gt_l = gt_l + far * (r(batch, 778, 3) * 0.03 + 0.02)
    gt_l[-1] += 1.0

    def perturb(v):
        a = r(batch, 3) * 0.25
        th = a.norm(dim=-1, keepdim=True)[..., None]
        k = a / a.norm(dim=-1, keepdim=True)
        K = torch.zeros(batch, 3, 3)
        K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -k[:, 2], k[:, 1], k[:, 2], -k[:, 0], -k[:, 1], k[:, 0]
        R = torch.eye(3) + torch.sin(th) * K + (1 - torch.cos(th)) * K.bmm(K)
        s = 1 + r(batch, 1, 1) * 0.08
        c = v.mean(1, keepdim=True)
        return s * (v - c).bmm(R.transpose(1, 2)) + c + r(batch, 1, 3) * 0.02 + r(batch, 778, 3) * 0.004
    return {'pred_left': perturb(gt_l), 'pred_right': perturb(gt_r), 'gt_left': gt_l, 'gt_right': gt_r}


def make_augment_case(n=3, size=256, seed=SEED):
    """Seeded loader inputs for the augmentation golden: smooth random uint8 BGR frames and a hand_dict shaped like the datasets' (core/loader.py:105-113)."""
    rng = np.random.RandomState(seed + 11)
    frames, dicts = [], []
    for i in range(n):
        f = rng.randint(0, 256, (size // 8, size // 8, 3)).astype(np.float32)
        f = np.kron(f, np.ones((8, 8, 1), np.float32))
        k = np.ones(5, np.float32) / 5                         # separable box blur: smooth gradients exercise the interpolation weights
        for ax in (0, 1):
            f = np.apply_along_axis(lambda m: np.convolve(m, k, mode='same'), ax, f)
        f = f + rng.randint(0, 40, f.shape)
        frames.append(np.clip(f, 0, 255).astype(np.uint8))
        hd = {}
        for side in ('left', 'right'):
            j3 = rng.randn(21, 3).astype(np.float32) * 0.04
            hd[side] = {'verts2d': (rng.rand(778, 2) * size).astype(np.float32), 'joints2d': (rng.rand(21, 2) * size).astype(np.float32),
                        'verts3d': rng.randn(778, 3).astype(np.float32) * 0.04, 'joints3d': j3}
        dicts.append(hd)
    return frames, dicts
