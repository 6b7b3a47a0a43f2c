"""Training losses of the hot path: `GraphLoss` + `calc_loss_GCN` (reference core/Loss.py:20-277) and, for the 'newgraph' MANO-tail
variant, `ManoLoss` + `mano_loss_GCN` (core/Loss_mano.py:62-335), same call signatures.

This is the *caller side* of the model.  On CUDA tensors the mesh terms of both hands (vertex, 2-D, joint,
face-normal, edge-length, coarse-level) run in the fused kernels of csrc/loss.cu (`rih_graph_loss_fwd/_bwd`: one forward and one backward
launch instead of ~400 small torch kernels); the handful of [B,48] / [B,10] pose and shape terms of `mano_loss_GCN` are torch ops on the
device.  The torch formulation in `GraphLoss` / `ManoLoss` is the readable statement of the same arithmetic: the host-side tests pin it to
the reference's loss code on CPU tensors, the GPU tests pin the fused kernels to it and to the fp64 oracle (`RIH_FUSED_LOSS=0` selects it
on the GPU for such A/B checks).
"""
import numpy as np
import torch
import torch.nn.functional as F


class GraphLoss:
    """core/Loss.py:20-175"""

    def __init__(self, J_regressor, faces, level=4, device='cuda', upsample_weight=None):
        self.device = device
        self.level = level + 1
        J = J_regressor.clone().detach().float()
        tips = torch.zeros_like(J[:5])
        for i, v in enumerate((745, 317, 444, 556, 673)):      # Loss.py:40-45
            tips[i, v] = 1.0
        J = torch.cat([J, tips], 0)
        order = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]
        self.J_regressor = J[order].contiguous().to(device)
        self.faces = torch.from_numpy(np.asarray(faces).astype(np.int64)).to(device)
        self.upsample_weight = None if upsample_weight is None else torch.as_tensor(upsample_weight).float().to(device)

    @staticmethod
    def _smooth_l1(a, b):
        return F.smooth_l1_loss(a, b)

    def mesh_downsample(self, feat, p=2):
        return F.avg_pool1d(feat.permute(0, 2, 1), p).permute(0, 2, 1)

    def _edges(self, v):
        e = v[:, self.faces]
        return torch.stack([e[:, :, 0] - e[:, :, 1], e[:, :, 1] - e[:, :, 2], e[:, :, 2] - e[:, :, 0]], 2)

    def norm_loss(self, verts_pred, verts_gt):
        eg, ep = self._edges(verts_gt), self._edges(verts_pred)
        n = F.normalize(torch.cross(eg[:, :, 0], eg[:, :, 1], dim=-1), dim=-1).unsqueeze(2)
        t = torch.sum(F.normalize(ep, dim=-1) * n, dim=-1)
        return self._smooth_l1(t, torch.zeros_like(t))

    def edge_loss(self, verts_pred, verts_gt):
        return self._smooth_l1(torch.linalg.norm(self._edges(verts_pred), dim=-1), torch.linalg.norm(self._edges(verts_gt), dim=-1))

    def calc_mano_loss(self, v3d_pred, v2d_pred, v3d_gt, v2d_gt, img_size):
        d = {}
        d['vert2d_loss'] = F.mse_loss(v2d_pred / img_size * 2 - 1, v2d_gt / img_size * 2 - 1)
        d['vert3d_loss'] = self._smooth_l1(v3d_pred, v3d_gt)
        d['joint_loss'] = self._smooth_l1(torch.matmul(self.J_regressor, v3d_pred), torch.matmul(self.J_regressor, v3d_gt))
        d['norm_loss'] = self.norm_loss(v3d_pred, v3d_gt)
        d['edge_loss'] = self.edge_loss(v3d_pred, v3d_gt)
        return d

    def upsample_weight_loss(self, w):
        x = w - self.upsample_weight
        return self._smooth_l1(x, torch.zeros_like(x))

    def calc_loss(self, converter, v3d_gt, v2d_gt, v3d_pred, v2d_pred, v3dList, v2dList, img_size):
        mano = self.calc_mano_loss(v3d_pred, v2d_pred, v3d_gt, v2d_gt, img_size)
        v3g, v2g = converter.vert_to_GCN(v3d_gt), converter.vert_to_GCN(v2d_gt)
        g3, g2 = [], []
        for _ in range(self.level):
            g3.append(v3g); g2.append(v2g)
            v3g, v2g = self.mesh_downsample(v3g), self.mesh_downsample(v2g)
        coarse = {'v3d_loss': [], 'v2d_loss': []}
        for a3, a2 in zip(v3dList, v2dList):
            j = [g.shape[1] for g in g3].index(a3.shape[1])
            coarse['v3d_loss'].append(self._smooth_l1(a3, g3[j]))
            coarse['v2d_loss'].append(F.mse_loss(a2 / img_size * 2 - 1, g2[j] / img_size * 2 - 1))
        return mano, coarse


class _FusedGraphLossFn(torch.autograd.Function):
    """rih_graph_loss_fwd / _bwd (csrc/loss.cu): both hands' GraphLoss.calc_loss + the calc_loss_GCN weighting as one forward and one
    backward kernel (the torch formulation above costs ~400 small launches, 2 ms of a 36 ms step)."""

    @staticmethod
    def forward(ctx, tables, weights, img, v3p_l, v2p_l, v3c_l, v2c_l, v3p_r, v2p_r, v3c_r, v2c_r, v3g_l, v2g_l, v3g_r, v2g_r, root_rel):
        import ctypes
        from ._lib import call
        from .ops import _stream
        dev = v3p_l.device
        if v3c_l is None:           # no coarse level (ManoLoss): zero-sized stand-ins keep the argument layout
            v3c_l = v3c_r = torch.empty(v3p_l.shape[0], 0, 3, device=dev)
            v2c_l = v2c_r = torch.empty(v3p_l.shape[0], 0, 2, device=dev)
        # the kernel hard-codes the MANO mesh (778 vertices) and a [B, V, 3|2] layout and reads raw pointers: check everything it will touch
        B0 = v3p_l.shape[0]
        Vc0 = v3c_l.shape[1]
        want = {'v3p_l': (B0, 778, 3), 'v2p_l': (B0, 778, 2), 'v3c_l': (B0, Vc0, 3), 'v2c_l': (B0, Vc0, 2),
                'v3p_r': (B0, 778, 3), 'v2p_r': (B0, 778, 2), 'v3c_r': (B0, Vc0, 3), 'v2c_r': (B0, Vc0, 2),
                'v3g_l': (B0, 778, 3), 'v2g_l': (B0, 778, 2), 'v3g_r': (B0, 778, 3), 'v2g_r': (B0, 778, 2), 'root_rel': (B0, 3)}
        got = dict(v3p_l=v3p_l, v2p_l=v2p_l, v3c_l=v3c_l, v2c_l=v2c_l, v3p_r=v3p_r, v2p_r=v2p_r, v3c_r=v3c_r, v2c_r=v2c_r,
                   v3g_l=v3g_l, v2g_l=v2g_l, v3g_r=v3g_r, v2g_r=v2g_r, root_rel=root_rel)
        for name, shape in want.items():
            t = got[name]
            if not (torch.is_tensor(t) and t.is_cuda and t.device == dev):
                raise RuntimeError('renderih_b200 fused loss: %s must be a CUDA tensor on %s (got %s); there is no CPU fallback'
                                   % (name, dev, getattr(t, 'device', type(t))))
            if tuple(t.shape) != shape:
                raise RuntimeError('renderih_b200 fused loss: %s has shape %s, expected %s' % (name, tuple(t.shape), shape))
        preds = [t.contiguous().float() for t in (v3p_l, v2p_l, v3c_l, v2c_l, v3p_r, v2p_r, v3c_r, v2c_r)]
        labels = [t.contiguous().float() for t in (v3g_l, v2g_l, v3g_r, v2g_r, root_rel)]
        B, Vc = preds[0].shape[0], preds[2].shape[1]
        F_ = tables['faces'][0].shape[0]
        pool = tables['perm'][0].numel() // Vc if Vc else 1
        fp = (ctypes.c_void_p * 16)(preds[0].data_ptr(), preds[1].data_ptr(), labels[0].data_ptr(), labels[1].data_ptr(), None,
                                    preds[2].data_ptr(), preds[3].data_ptr(), tables['J21'][0].data_ptr(),
                                    preds[4].data_ptr(), preds[5].data_ptr(), labels[2].data_ptr(), labels[3].data_ptr(), labels[4].data_ptr(),
                                    preds[6].data_ptr(), preds[7].data_ptr(), tables['J21'][1].data_ptr())
        ip = (ctypes.c_void_p * 4)(tables['faces'][0].data_ptr(), tables['perm'][0].data_ptr(), tables['faces'][1].data_ptr(), tables['perm'][1].data_ptr())
        partial = torch.empty((2, B, 7), device=dev)
        out = torch.empty(15, device=dev)
        coef = torch.empty(7, device=dev)
        w = (ctypes.c_float * 7)(*weights)
        call('rih_graph_loss_fwd', fp, ip, B, F_, Vc, pool, float(img), w, partial.data_ptr(), out.data_ptr(), coef.data_ptr(), _stream())
        ctx.save_for_backward(coef, *preds, *labels)
        ctx.meta = (tables, B, F_, Vc, pool, float(img))
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, d_total, _d_out):
        import ctypes
        from ._lib import call
        from .ops import _stream
        coef, *rest = ctx.saved_tensors
        preds, labels = rest[:8], rest[8:]
        tables, B, F_, Vc, pool, img = ctx.meta
        fp = (ctypes.c_void_p * 16)(preds[0].data_ptr(), preds[1].data_ptr(), labels[0].data_ptr(), labels[1].data_ptr(), None,
                                    preds[2].data_ptr(), preds[3].data_ptr(), tables['J21'][0].data_ptr(),
                                    preds[4].data_ptr(), preds[5].data_ptr(), labels[2].data_ptr(), labels[3].data_ptr(), labels[4].data_ptr(),
                                    preds[6].data_ptr(), preds[7].data_ptr(), tables['J21'][1].data_ptr())
        ip = (ctypes.c_void_p * 4)(tables['faces'][0].data_ptr(), tables['perm'][0].data_ptr(), tables['faces'][1].data_ptr(), tables['perm'][1].data_ptr())
        grads = [torch.empty_like(t) for t in preds]
        gp = (ctypes.c_void_p * 8)(*[g.data_ptr() for g in grads])
        up = d_total.contiguous().float().reshape(1)
        scratch = torch.empty(7, device=up.device)
        call('rih_graph_loss_bwd', fp, ip, gp, B, F_, Vc, pool, img, coef.data_ptr(), up.data_ptr(), scratch.data_ptr(), _stream())
        grads = [g if g.numel() else None for g in grads]
        return (None, None, None) + tuple(grads) + (None, None, None, None, None)


def _fused_tables(gl_left, gl_right, conv_left, conv_right, device):
    """int32 face / permutation tables and the 21-joint regressors on the device (cached on the GraphLoss objects)."""
    key = str(device)
    cache = gl_left.__dict__.setdefault('_fused_tables', {})
    if key not in cache:
        t = {'faces': [], 'perm': [], 'J21': []}
        for gl, conv in ((gl_left, conv_left), (gl_right, conv_right)):
            t['faces'].append(gl.faces.to(device=device, dtype=torch.int32).contiguous())
            perm = np.asarray(conv.graph_perm if conv is not None else [0], dtype=np.int32)      # ManoLoss: no coarse level, no permutation
            if perm.min() < 0 or perm.max() >= 778 or int(gl.faces.max()) >= 778:
                raise ValueError('fused GraphLoss: graph_perm / faces index outside the 778 MANO vertices')
            t['perm'].append(torch.as_tensor(perm).to(device))
            t['J21'].append(gl.J_regressor.to(device=device, dtype=torch.float32).contiguous())
        cache[key] = t
    return cache[key]


def calc_loss_GCN(cfg, epoch, graph_loss_left, graph_loss_right, converter_left, converter_right,
                  result, paramsDict, handDictList, otherInfo, mask, dense, hms,
                  v2d_l, j2d_l, v2d_r, j2d_r, v3d_l, j3d_l, v3d_r, j3d_r, root_rel, img_size, upsample_weight=None):
    """core/Loss.py:201-277 (the auxiliary mask/dense/heat-map loss is disabled in the reference at line 213).
    CUDA inputs with one coarse level (the models of this package) take the fused kernel; anything else the torch formulation below."""
    import os
    v3p = result['verts3d']['left']
    if (v3p.is_cuda and len(handDictList) == 1 and type(graph_loss_left) is GraphLoss and os.environ.get('RIH_FUSED_LOSS', '1') != '0'
            and handDictList[0]['verts3d']['left'].shape[1] * 4 == len(converter_left.graph_perm)):
        w = cfg.LOSS_WEIGHT
        alpha = 0 if epoch < w.GRAPH.NORM.NORM_EPOCH else 1
        weights = (w.DATA.LABEL_3D, w.DATA.LABEL_2D, w.DATA.LABEL_3D, w.GRAPH.NORM.NORMAL, alpha * w.GRAPH.NORM.EDGE, w.DATA.LABEL_3D, w.DATA.LABEL_2D)
        tables = _fused_tables(graph_loss_left, graph_loss_right, converter_left, converter_right, v3p.device)
        hd = handDictList[0]
        total, out = _FusedGraphLossFn.apply(tables, weights, img_size, result['verts3d']['left'], result['verts2d']['left'],
                                             hd['verts3d']['left'], hd['verts2d']['left'], result['verts3d']['right'], result['verts2d']['right'],
                                             hd['verts3d']['right'], hd['verts2d']['right'], v3d_l, v2d_l, v3d_r, v2d_r, root_rel)
        names = ('vert3d_loss', 'vert2d_loss', 'joint_loss', 'norm_loss', 'edge_loss')
        mano = {n: (out[1 + i] + out[8 + i]) / 2 for i, n in enumerate(names)}
        coarse = {'v3d_loss': [(out[6] + out[13]) / 2], 'v2d_loss': [(out[7] + out[14]) / 2]}
        if upsample_weight is not None:
            mano['upsample_norm_loss'] = graph_loss_left.upsample_weight_loss(upsample_weight)
            total = total + w.NORM.UPSAMPLE * mano['upsample_norm_loss']
        else:
            mano['upsample_norm_loss'] = torch.zeros_like(total)
        return total, {'total_loss': 0}, mano, coarse
    aux = {'total_loss': 0}
    v3d_r = v3d_r + root_rel.unsqueeze(1)
    outs = {}
    for side, gl, conv, v3g, v2g in (('left', graph_loss_left, converter_left, v3d_l, v2d_l),
                                     ('right', graph_loss_right, converter_right, v3d_r, v2d_r)):
        outs[side] = gl.calc_loss(conv, v3g, v2g, result['verts3d'][side], result['verts2d'][side],
                                  [h['verts3d'][side] for h in handDictList], [h['verts2d'][side] for h in handDictList], img_size)
    mano = {k: (outs['left'][0][k] + outs['right'][0][k]) / 2 for k in outs['left'][0]}
    coarse = {k: [(a + b) / 2 for a, b in zip(outs['left'][1][k], outs['right'][1][k])] for k in outs['left'][1]}
    w = cfg.LOSS_WEIGHT
    alpha = 0 if epoch < w.GRAPH.NORM.NORM_EPOCH else 1
    if upsample_weight is not None:
        mano['upsample_norm_loss'] = graph_loss_left.upsample_weight_loss(upsample_weight)
    else:
        mano['upsample_norm_loss'] = torch.zeros_like(mano['vert3d_loss'])
    total = w.DATA.LABEL_3D * mano['vert3d_loss'] + w.DATA.LABEL_2D * mano['vert2d_loss'] + w.DATA.LABEL_3D * mano['joint_loss'] \
        + w.GRAPH.NORM.NORMAL * mano['norm_loss'] + alpha * w.GRAPH.NORM.EDGE * mano['edge_loss']
    for i in range(len(coarse['v3d_loss'])):
        total = total + w.DATA.LABEL_3D * coarse['v3d_loss'][i] + w.DATA.LABEL_2D * coarse['v2d_loss'][i]
    total = total + w.NORM.UPSAMPLE * mano['upsample_norm_loss']
    return total, aux, mano, coarse


# ----------------------------------------------------------------------------- 'newgraph' variant: core/Loss_mano.py
def axis_angle_to_rotmat_quat(axisang):
    """Axis-angle [N,3] -> rotation matrices [N,9] through a unit quaternion -- the construction `batch_rodrigues` of core/Loss_mano.py:48-60
    uses for the pose term (note its 1e-8 is added to the vector BEFORE the norm, unlike manolayer.rodrigues_batch)."""
    angle = torch.norm(axisang + 1e-8, p=2, dim=1, keepdim=True)
    q = torch.cat([torch.cos(angle * 0.5), torch.sin(angle * 0.5) * (axisang / angle)], dim=1)
    q = q / q.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([w * w + x * x - y * y - z * z, 2 * x * y - 2 * w * z, 2 * w * y + 2 * x * z,
                        2 * w * z + 2 * x * y, w * w - x * x + y * y - z * z, 2 * y * z - 2 * w * x,
                        2 * x * z - 2 * w * y, 2 * w * x + 2 * y * z, w * w - x * x - y * y + z * z], dim=1)


class ManoLoss(GraphLoss):
    """core/Loss_mano.py:62-215: the GraphLoss mesh terms plus MSE on the 16 joint rotation matrices and on the shape coefficients;
    no coarse-level terms (they are commented out in the reference, :184-208)."""

    def calc_mano_loss(self, v3d_pred, v2d_pred, v3d_gt, v2d_gt, img_size, pred_pose, pred_shape, pose_gt, shape_gt):
        d = GraphLoss.calc_mano_loss(self, v3d_pred, v2d_pred, v3d_gt, v2d_gt, img_size)
        d['pose_loss'] = F.mse_loss(axis_angle_to_rotmat_quat(pred_pose.reshape(-1, 3)).reshape(-1, 16, 3, 3),
                                    axis_angle_to_rotmat_quat(pose_gt.reshape(-1, 3)).reshape(-1, 16, 3, 3))
        d['shape_loss'] = F.mse_loss(pred_shape, shape_gt)
        return d

    def calc_loss(self, converter, v3d_gt, v2d_gt, v3d_pred, v2d_pred, v3dList, v2dList, img_size, pred_pose, pred_shape, pose_gt, shape_gt):
        return self.calc_mano_loss(v3d_pred, v2d_pred, v3d_gt, v2d_gt, img_size, pred_pose, pred_shape, pose_gt, shape_gt)


def mano_loss_GCN(cfg, epoch, graph_loss_left, graph_loss_right, converter_left, converter_right,
                  result, paramsDict, handDictList, otherInfo, mask, dense, hms,
                  v2d_l, j2d_l, v2d_r, j2d_r, v3d_l, j3d_l, v3d_r, j3d_r, root_rel, img_size, lp_gt, ls_gt, rp_gt, rs_gt, upsample_weight=None):
    """core/Loss_mano.py:245-335: mesh terms on the MANO vertices, pose / shape / relative-root terms, shape regulariser."""
    import os
    aux = {'total_loss': 0}
    ml = otherInfo['verts3d_MANO_list']
    if result['verts3d']['left'].is_cuda and type(graph_loss_left) is ManoLoss and os.environ.get('RIH_FUSED_LOSS', '1') != '0':
        # mesh terms of both hands (vert3d, vert2d, joint, normal, edge) from the fused kernel; the small pose / shape / root terms stay torch ops
        w = cfg.LOSS_WEIGHT
        alpha = 0 if epoch < w.GRAPH.NORM.NORM_EPOCH else 1
        weights = (w.DATA.LABEL_3D, w.DATA.LABEL_2D, w.DATA.LABEL_3D, w.GRAPH.NORM.NORMAL, alpha * w.GRAPH.NORM.EDGE, 0.0, 0.0)
        tables = _fused_tables(graph_loss_left, graph_loss_right, None, None, result['verts3d']['left'].device)
        total, out = _FusedGraphLossFn.apply(tables, weights, img_size, result['verts3d']['left'], result['verts2d']['left'], None, None,
                                             result['verts3d']['right'], result['verts2d']['right'], None, None, v3d_l, v2d_l, v3d_r, v2d_r, root_rel)
        mano = {n: (out[1 + i] + out[8 + i]) / 2 for i, n in enumerate(('vert3d_loss', 'vert2d_loss', 'joint_loss', 'norm_loss', 'edge_loss'))}
        rot = lambda p: axis_angle_to_rotmat_quat(p.reshape(-1, 3)).reshape(-1, 16, 3, 3)
        mano['pose_loss'] = (F.mse_loss(rot(ml['left']['mano_pose']), rot(lp_gt)) + F.mse_loss(rot(ml['right']['mano_pose']), rot(rp_gt))) / 2
        mano['shape_loss'] = (F.mse_loss(ml['left']['mano_shape'], ls_gt) + F.mse_loss(ml['right']['mano_shape'], rs_gt)) / 2
        if upsample_weight is not None:
            mano['upsample_norm_loss'] = graph_loss_left.upsample_weight_loss(upsample_weight)
        else:
            mano['upsample_norm_loss'] = torch.zeros_like(total)
        mano['rootrel_loss'] = w.DATA.MANO_REL * F.mse_loss(otherInfo['root_rel'], root_rel)
        mano['regularize_loss'] = 0.005 * torch.mean(torch.sum(ml['left']['mano_shape'] ** 2) + torch.sum(ml['right']['mano_shape'] ** 2))
        total = total + w.DATA.MANO_POSE * mano['pose_loss'] + w.DATA.MANO_SHAPE * mano['shape_loss'] + mano['rootrel_loss'] + mano['regularize_loss'] \
            + w.NORM.UPSAMPLE * mano['upsample_norm_loss']
        return total, aux, mano, {}
    v3d_r = v3d_r + root_rel.unsqueeze(1)
    left = graph_loss_left.calc_loss(converter_left, v3d_l, v2d_l, result['verts3d']['left'], result['verts2d']['left'], None, None, img_size,
                                     ml['left']['mano_pose'], ml['left']['mano_shape'], lp_gt, ls_gt)
    right = graph_loss_right.calc_loss(converter_right, v3d_r, v2d_r, result['verts3d']['right'], result['verts2d']['right'], None, None, img_size,
                                       ml['right']['mano_pose'], ml['right']['mano_shape'], rp_gt, rs_gt)
    mano = {k: (left[k] + right[k]) / 2 for k in left}
    w = cfg.LOSS_WEIGHT
    alpha = 0 if epoch < w.GRAPH.NORM.NORM_EPOCH else 1
    if upsample_weight is not None:
        mano['upsample_norm_loss'] = graph_loss_left.upsample_weight_loss(upsample_weight)
    else:
        mano['upsample_norm_loss'] = torch.zeros_like(mano['vert3d_loss'])
    mano['rootrel_loss'] = w.DATA.MANO_REL * F.mse_loss(otherInfo['root_rel'], root_rel)
    mano['regularize_loss'] = 0.005 * torch.mean(torch.sum(ml['left']['mano_shape'] ** 2) + torch.sum(ml['right']['mano_shape'] ** 2))
    total = w.DATA.LABEL_3D * mano['vert3d_loss'] + w.DATA.LABEL_2D * mano['vert2d_loss'] + w.DATA.LABEL_3D * mano['joint_loss'] \
        + w.GRAPH.NORM.NORMAL * mano['norm_loss'] + alpha * w.GRAPH.NORM.EDGE * mano['edge_loss'] \
        + w.DATA.MANO_POSE * mano['pose_loss'] + w.DATA.MANO_SHAPE * mano['shape_loss'] + mano['rootrel_loss'] + mano['regularize_loss']
    total = total + w.NORM.UPSAMPLE * mano['upsample_norm_loss']
    return total, aux, mano, {}
