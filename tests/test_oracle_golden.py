"""Pin the CPU oracle (oracle/model_ref.py, oracle/mano_ref.py) to the golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py).  CPU only.  Tolerances are float32 round-off of differently associated ops."""
import os

import numpy as np
import pytest
import torch

from oracle import fixtures, mano_ref, model_ref
from renderih_b200 import assets as rih_assets

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
RTOL = 2e-5   # relative to the tensor's max magnitude


def rel_err(a, b):
    a, b = torch.as_tensor(a).detach().double(), torch.as_tensor(b).detach().double()
    return float(((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).detach())


# (golden file, first BatchNorm of the trunk) per model: ResNet-50 (BASELINE configs 2-4), HRNet-w48 (config 5), myhand graph variant
CASES = {'resnet50': ('model_synth_b2.pt', 'encoder.resnet.bn1'), 'hrnet48': ('model_hrnet48_synth_b2.pt', 'encoder.hrnet.bn1'),
         'graph': ('model_graph_synth_b2.pt', 'encoder.resnet.bn1'),      # 'graph' = common/myhand default variant
         'resnet50_b16': ('model_synth_b16.pt', 'encoder.resnet.bn1')}    # batch 16: well-conditioned train-mode BatchNorm (>= 1024 samples per channel)


# hrnet_mid's biased convolutions that feed a BatchNorm (models/encoder.py:304-326)
BN_SHADOWED_BIAS = {'mid_model.downsamp_modules.%d.0.bias' % i for i in range(3)} | {'mid_model.final_layer.0.bias'}


@pytest.fixture(scope='module', params=sorted(CASES))
def gold(request):
    g = torch.load(os.path.join(GOLD, CASES[request.param][0]), weights_only=False)
    g['encoder_type'] = request.param.split('_')[0]
    g['case'] = request.param
    return g


@pytest.fixture(scope='module')
def setup(gold):
    from renderih_b200.config import load_cfg
    from renderih_b200.model import load_model
    a = rih_assets.synthetic_assets(0)
    cfg = load_cfg(None)
    if gold['encoder_type'] == 'graph':
        from renderih_b200.myhand import load_graph_model
        tmpl = load_graph_model(cfg, assets=a, mano_assets={s_: rih_assets.synthetic_mano(0, s_) for s_ in ('left', 'right')}).state_dict()
    else:
        cfg.MODEL.ENCODER_TYPE = gold['encoder_type']
        tmpl = load_model(cfg, assets=a).state_dict()
    sd = fixtures.init_state_dict(tmpl)
    assert fixtures.checksum(sd) == gold['weights_sha256'], 'deterministic weight init drifted from the golden run'
    return a, sd


def flat(out):
    result, params, hlist, other = out
    d = {}
    for side in ('left', 'right'):
        d['verts3d_' + side] = result['verts3d'][side]; d['verts2d_' + side] = result['verts2d'][side]
        d['scale_' + side] = params['scale'][side]; d['trans2d_' + side] = params['trans2d'][side]
        d['v3c_' + side] = hlist[0]['verts3d'][side]; d['v2c_' + side] = hlist[0]['verts2d'][side]
        if other['verts3d_MANO_list'][side]:
            d['v3list_' + side] = other['verts3d_MANO_list'][side][0]; d['v2list_' + side] = other['verts2d_MANO_list'][side][0]
    for k in (('hms', 'mask', 'dense') if 'hms' in other else ()):
        t = other[k] if other[k].dim() == 4 else other[k][:, None]      # HRnet_encoder's mask is [B,64,64] (models/encoder.py:235)
        d[k + '_sub'] = t[:, :, ::8, ::8]; d[k + '_mean'] = t.mean(dim=(2, 3))
    return d


def test_oracle_eval_forward_matches_reference_golden(gold, setup):
    a, sd = setup
    sd = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        out = flat(model_ref.model_forward(sd, model_ref.prepare_assets(a), fixtures.make_image(gold['batch']), training=False))
    for k, v in gold['eval'].items():
        assert out[k].shape == v.shape, k
        assert rel_err(out[k], v) < RTOL, (k, rel_err(out[k], v))


def test_oracle_train_forward_backward_matches_reference_golden(gold, setup):
    a, sd = setup
    sd = {k: v.clone() for k, v in sd.items()}
    for k, v in sd.items():
        if v.is_floating_point() and 'running_' not in k and '.mano_' not in k and k not in ('decoder.dense_coor', 'decoder.unsample_layer.weight'):
            v.requires_grad_(True)
    out = model_ref.model_forward(sd, model_ref.prepare_assets(a), fixtures.make_image(gold['batch']), training=True, dropout=0.0)
    fo = flat(out)
    for k, v in gold['train']['out'].items():
        assert rel_err(fo[k], v) < RTOL, (k, rel_err(fo[k], v))
    la = fixtures.make_loss_assets(a, rih_assets.synthetic_mano(0, 'left'), rih_assets.synthetic_mano(0, 'right'))
    loss = model_ref.calc_loss_GCN(out, fixtures.make_labels(gold['batch']), la)
    assert abs(float(loss) - gold['train']['loss']) / gold['train']['loss'] < 1e-5
    loss.backward()
    bn1 = CASES[gold['case']][1]
    assert rel_err(sd[bn1 + '.running_mean'], gold['train']['bn1_running_mean']) < RTOL
    assert rel_err(sd[bn1 + '.running_var'], gold['train']['bn1_running_var']) < RTOL
    for k in gold['train']['no_grad_keys']:
        g = sd[k].grad
        assert g is None or float(g.abs().max()) == 0.0, k
    worst = 0.0
    for k, g in gold['train']['grads'].items():
        mine = sd[k].grad
        assert mine is not None, k
        if k.endswith('w_ks.bias'):
            # a key bias shifts every score of a query row equally: softmax-invariant, the true gradient is 0
            # and both sides hold only round-off noise
            assert float(mine.norm()) < 1e-3 * (1.0 + float(sd[k.replace('w_ks.bias', 'w_qs.bias')].grad.norm())), k
            continue
        if k in BN_SHADOWED_BIAS:
            # a convolution bias directly in front of a train-mode BatchNorm is removed by the mean subtraction: the true gradient
            # is 0, both sides hold round-off noise (tiny next to the gradient of the convolution's weight)
            assert float(mine.norm()) < 1e-3 * float(sd[k.replace('.bias', '.weight')].grad.norm()), k
            continue
        e = abs(float(mine.norm()) - g['norm']) / max(g['norm'], 1e-6)
        worst = max(worst, e)
        assert e < 2e-3, (k, e)
        if 'full' in g and g['norm'] > 1e-3:
            assert rel_err(mine, g['full']) < 2e-3, k


def test_mano_oracle_matches_reference_golden():
    mg = torch.load(os.path.join(GOLD, 'mano_synth.pt'), weights_only=False)
    inp = fixtures.make_mano_inputs(5)
    root = mano_ref.rodrigues(inp['axis'].numpy())
    assert np.abs(root - mg['rodrigues'].numpy()).max() < 1e-6
    for case in mg['cases']:
        c = case['cfg']
        m = rih_assets.synthetic_mano(0, case['side'])
        m = dict(m); m['J_regressor'] = np.asarray(m['J_regressor'].todense())
        if c['use_pca']:
            pose = inp['pose_pca'][:, :c['ncomps']].numpy()
        else:
            pose = mano_ref.rodrigues(inp['pose_axis'].numpy().reshape(-1, 3)).reshape(-1, 15, 3, 3)
        tr, sc = (inp['trans'].numpy(), inp['scale'].numpy()) if c['ts'] else (None, None)
        v, j = mano_ref.mano_forward(m, root, pose, inp['shape'].numpy(), tr, sc, use_pca=c['use_pca'],
                                     center_idx=c['center_idx'], new_skel=c['new_skel'])
        assert np.abs(v - case['v'].numpy()).max() < 2e-6, case['cfg']
        assert np.abs(j - case['j'].numpy()).max() < 2e-6, case['cfg']


def _mano_case_inputs(c, layer_axis2Rmat=None):
    inp = fixtures.make_mano_inputs(5)
    root = torch.from_numpy(mano_ref.rodrigues(inp['axis'].numpy()))
    if c['use_pca']:
        pose = inp['pose_pca'][:, :c['ncomps']].clone()
    else:
        pose = torch.from_numpy(mano_ref.rodrigues(inp['pose_axis'].numpy().reshape(-1, 3)).reshape(-1, 15, 3, 3))
    tr, sc = (inp['trans'].clone(), inp['scale'].clone()) if c['ts'] else (None, None)
    return root, pose, inp['shape'].clone(), tr, sc


def test_mano_torch_oracle_gradients_match_reference_golden():
    """The differentiable restatement (oracle/mano_ref.mano_forward_torch) against the UNMODIFIED reference ManoLayer's autograd
    gradients (tests/golden/mano_grad_synth.pt): 2e-5 relative to each gradient tensor's max."""
    gg = torch.load(os.path.join(GOLD, 'mano_grad_synth.pt'), weights_only=False)
    wv, wj = fixtures.make_mano_loss_weights(5)
    for case in gg['cases']:
        c = case['cfg']
        m = rih_assets.synthetic_mano(0, case['side'])
        leaves = [None if t is None else t.requires_grad_(True) for t in _mano_case_inputs(c)]
        v, j = mano_ref.mano_forward_torch(m, *leaves, use_pca=c['use_pca'], center_idx=c['center_idx'], new_skel=c['new_skel'])
        vn, jn = mano_ref.mano_forward(dict(m, J_regressor=np.asarray(m['J_regressor'].todense())), leaves[0].detach().numpy(), leaves[1].detach().numpy(),
                                       leaves[2].detach().numpy(), None if leaves[3] is None else leaves[3].detach().numpy(),
                                       None if leaves[4] is None else leaves[4].detach().numpy(), use_pca=c['use_pca'], center_idx=c['center_idx'], new_skel=c['new_skel'])
        assert np.abs(v.detach().numpy() - vn).max() < 2e-6 and np.abs(j.detach().numpy() - jn).max() < 2e-6
        ((v * wv).sum() + (j * wj).sum()).backward()
        for name, t in zip(('d_root', 'd_pose', 'd_shape', 'd_trans', 'd_scale'), leaves):
            if t is None:
                assert case[name] is None
                continue
            assert rel_err(t.grad, case[name]) < 2e-5, (case['side'], c, name, rel_err(t.grad, case[name]))


# ----------------------------------------------------------------------------- 'newgraph' variant (graph decoder + ParamRegressor + MANO tail)
@pytest.fixture(scope='module')
def newgraph_setup():
    from renderih_b200.myhand import load_new_model
    gold = torch.load(os.path.join(GOLD, 'model_newgraph_synth_b2.pt'), weights_only=False)
    a = rih_assets.synthetic_assets(0)
    manos = {s_: rih_assets.synthetic_mano(0, s_) for s_ in ('left', 'right')}
    tmpl = load_new_model(None, assets=a, mano_assets=manos).state_dict()
    sd = fixtures.init_state_dict(tmpl)
    assert len(sd) == 1063 and fixtures.checksum(sd) == gold['weights_sha256'], 'state_dict keys / order / deterministic init drifted from the reference'
    A = fixtures.add_mano_assets(model_ref.prepare_assets(a), manos['left'], manos['right'])
    return gold, sd, A


def test_newgraph_oracle_matches_reference_golden(newgraph_setup):
    """oracle/model_ref.newgraph_tail (+ the differentiable MANO restatement) against the unmodified reference's load_new_model outputs
    and gradients (built on the CPU through oracle/ref_bridge shims): forward 2e-5, loss 1e-5, gradient norms 2e-3."""
    gold, sd0, A = newgraph_setup
    with torch.no_grad():
        out = fixtures.flat_newgraph(model_ref.model_forward({k: v.clone() for k, v in sd0.items()}, A, fixtures.make_image(2), training=False))
    for k, v in gold['eval'].items():
        assert out[k].shape == v.shape and rel_err(out[k], v) < RTOL, (k, rel_err(out[k], v))
    sd = {k: v.clone() for k, v in sd0.items()}
    for k, v in sd.items():
        if v.is_floating_point() and 'running_' not in k and '.mano_' not in k and k not in ('decoder.dense_coor', 'decoder.unsample_layer.weight'):
            v.requires_grad_(True)
    fo = fixtures.flat_newgraph(model_ref.model_forward(sd, A, fixtures.make_image(2), training=True, dropout=0.0))
    for k, v in gold['train']['out'].items():
        assert rel_err(fo[k], v) < RTOL, (k, rel_err(fo[k], v))
    loss = fixtures.newgraph_loss(fo, fixtures.make_newgraph_cotangents(2))
    assert abs(float(loss) - gold['train']['loss']) < 1e-5 * max(1.0, abs(gold['train']['loss'])) + 2e-3
    loss.backward()
    for k in gold['train']['no_grad_keys']:
        assert sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0, k
    for k, g in gold['train']['grads'].items():
        mine = sd[k].grad
        assert mine is not None, k
        if k.endswith('w_ks.bias'):
            continue
        assert abs(float(mine.norm()) - g['norm']) / max(g['norm'], 1e-6) < 2e-3, (k, float(mine.norm()), g['norm'])


def test_eval_metrics_oracle_matches_reference_golden():
    """oracle/metrics_ref.py against the numbers produced by executing apps/eval_interhand.py's own loop body and reductions
    (tests/golden/eval_metrics_synth.pt, two batches of 4 and 2 samples; the 2-sample batch exercises the reference's layout quirk)."""
    import numpy as np
    from oracle import metrics_ref
    from renderih_b200 import assets as A
    gold = torch.load(os.path.join(GOLD, 'eval_metrics_synth.pt'), weights_only=False)
    case = fixtures.make_eval_case(6)
    J = {s: metrics_ref.joint_regressor21(torch.from_numpy(np.asarray(A.synthetic_mano(0, s)['J_regressor'].todense(), dtype='float32'))) for s in ('left', 'right')}
    outs = [metrics_ref.eval_batch(J['left'], J['right'], *[case[k][lo:hi] for k in ('pred_left', 'pred_right', 'gt_left', 'gt_right')])
            for lo, hi in ((0, 4), (4, 6))]
    for side in ('left', 'right'):
        for k, v in gold['per_element'][side].items():
            o = torch.cat([b[side][k] for b in outs])
            assert o.shape == v.shape
            assert float((o - v).abs().max()) <= 1e-6 * float(v.abs().max()), (side, k)
    assert torch.equal(torch.cat([b['mrrpe'] for b in outs]), gold['mrrpe'])
    cd = torch.cat([b['cdev'] for b in outs])
    assert torch.isnan(cd[-1]) and torch.isnan(gold['cdev'][-1])
    assert torch.allclose(cd[:-1], gold['cdev'][:-1], rtol=1e-6)
    for k, name in (('double_pa_joint', 'pa_joint'), ('double_pa_mesh', 'pa_mesh')):
        assert np.allclose(np.concatenate([b[k] for b in outs]), gold['double_per_sample'][name].numpy(), rtol=1e-6)
    s = metrics_ref.summarize(outs)
    g = gold['summary_mm']
    for side in ('left', 'right'):
        for k in ('ori_mpjpe', 'ori_mpvpe', 'mpjpe', 'mpvpe', 'pa_mpjpe', 'pa_mpvpe'):
            assert abs(s[side][k] - g['%s_%s' % (k, side)]) <= 1e-6 * g['%s_%s' % (k, side)], (side, k)
    for k, gk in (('double_pa_joint', 'double_pa_mpjpe'), ('double_pa_mesh', 'double_pa_mpvpe'), ('double_joint', 'double_mpjpe'), ('double_mesh', 'double_mpvpe')):
        assert abs(s[k] - g[gk]) <= 1e-6 * g[gk], k
    # without the reference's batch-size quirk the 4-sample batch is unchanged and the 2-sample one differs (documented deviation of the kernel)
    plain = metrics_ref.eval_batch(J['left'], J['right'], *[case[k][4:6] for k in ('pred_left', 'pred_right', 'gt_left', 'gt_right')], batch_quirk=False)
    assert float((plain['left']['pajoints_loss'] - outs[1]['left']['pajoints_loss']).abs().max()) > 1e-4
    plain4 = metrics_ref.eval_batch(J['left'], J['right'], *[case[k][0:4] for k in ('pred_left', 'pred_right', 'gt_left', 'gt_right')], batch_quirk=False)
    assert torch.equal(plain4['left']['pajoints_loss'], outs[0]['left']['pajoints_loss'])


def test_augment_oracle_matches_reference_golden_and_cv2():
    """oracle/augment_ref.py (fixed-point restatement of cv2.warpAffine + the loader's noise / flip / normalise / label steps) against
    tests/golden/augment_synth.pt = outputs of the reference's own handDataset.process_data, bit-exact for the images; and, when cv2 is
    importable, against cv2.warpAffine directly over a sweep of the loader's augmentation ranges and some extreme maps."""
    import numpy as np
    from oracle import augment_ref as ar
    gold = torch.load(os.path.join(GOLD, 'augment_synth.pt'), weights_only=False)
    frames, dicts = fixtures.make_augment_case(3)
    for i, s in enumerate(gold['samples']):
        img, ori, net, M = ar.process_image(frames[i], s['theta'], s['scale'], s['u'], s['v'], s['a'], s['b'], s['flip'])
        final = img[:, ::-1] if s['flip'] else img
        assert np.array_equal(final, s['final_u8_bgr'].numpy()), i
        if s['imgTensor'] is not None:
            assert np.array_equal(net, s['imgTensor'].numpy())
        lab = ar.process_labels(dicts[i], s['theta'], M, s['flip'], bone_length=gold['meta']['bone_length'])
        for k, v in s['labels'].items():
            assert np.array_equal(lab[k], v.numpy()), (i, k)
    try:
        import cv2
    except ImportError:
        return
    rng = np.random.RandomState(5)
    cases = [(0, 1, 0, 0), (90, 1, 0, 0), (-90, 0.5, 100, -100), (45, 3.0, 0, 0), (180, 1, 0.5, 0.5), (0, 1, 300, 0)]
    cases += [(rng.uniform(-90, 90), rng.uniform(0.75, 1.25), rng.uniform(-10, 10), rng.uniform(-10, 10)) for _ in range(12)]
    for j, (theta, sc, u, v) in enumerate(cases):
        img = frames[j % 3] if j % 2 else rng.randint(0, 256, (256, 256, 3)).astype(np.uint8)
        M = ar.get_affine_mat(theta, sc, u, v, 256, 256)
        assert np.array_equal(cv2.warpAffine(src=img, M=M[0:2, :], dsize=(256, 256)), ar.warp_affine_u8(img, M[0:2, :])), (theta, sc, u, v)
