"""CPU oracle: numpy restatement of `ManoLayer.forward` (reference models/manolayer.py:250-322).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Written as explicit per-step numpy (float32) following the
reference order of operations; pinned against the unmodified reference by oracle/make_golden.py.
"""
import numpy as np

NEW_ORDER = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]   # manolayer.py:110-115
TIPS = [745, 317, 444, 556, 673]                                                         # manolayer.py:296
PARENT = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]


def rodrigues(axis):
    """rodrigues_batch, manolayer.py:32-48 (eps added AFTER the norm, line 37)"""
    axis = np.asarray(axis, np.float32)
    angle = (np.linalg.norm(axis, axis=1, keepdims=True) + np.float32(1e-8)).astype(np.float32)
    a = axis / angle
    s, c = np.sin(angle)[..., None], np.cos(angle)[..., None]
    L = np.zeros((axis.shape[0], 3, 3), np.float32)
    L[:, 2, 1] = a[:, 0]; L[:, 1, 2] = -a[:, 0]
    L[:, 0, 2] = a[:, 1]; L[:, 2, 0] = -a[:, 1]
    L[:, 1, 0] = a[:, 2]; L[:, 0, 1] = -a[:, 2]
    return (np.eye(3, dtype=np.float32)[None] + s * L + (1 - c) * (L @ L)).astype(np.float32)


def mano_forward(m, root_rotation, pose, shape, trans=None, scale=None, use_pca=True, center_idx=9, new_skel=False, parent=PARENT):
    """m: dict with hands_components, hands_mean, shapedirs[778,3,10], posedirs[778,3,135], v_template, J_regressor[16,778] dense, weights."""
    f = lambda a: np.asarray(a, np.float32)
    R0, pose, shape = f(root_rotation), f(pose), f(shape)
    bs = R0.shape[0]
    if use_pca:   # pca2Rmat, manolayer.py:163-175
        axis = pose @ f(m['hands_components'])[:pose.shape[1]] + f(m['hands_mean'])
        rot = rodrigues(axis.reshape(-1, 3)).reshape(bs, 15, 3, 3)
    else:
        rot = pose.reshape(bs, 15, 3, 3)
    v_shaped = f(m['v_template'])[None] + np.einsum('vck,bk->bvc', f(m['shapedirs']), shape)      # :264-265
    j_tpose = np.einsum('jv,bvc->bjc', f(m['J_regressor']), v_shaped)                              # :267
    pose_shape = (rot - np.eye(3, dtype=np.float32)).reshape(bs, 135)                              # :269-270
    v_tpose = v_shaped + np.einsum('vck,bk->bvc', f(m['posedirs']), pose_shape)                    # :271-272
    SE3 = np.zeros((bs, 16, 4, 4), np.float32)
    for i in range(16):                                                                            # :274-283
        R = R0 if i == 0 else rot[:, i - 1]
        t = np.einsum('bij,bj->bi', np.eye(3, dtype=np.float32)[None] - R, j_tpose[:, i])
        loc = np.zeros((bs, 4, 4), np.float32); loc[:, :3, :3] = R; loc[:, :3, 3] = t; loc[:, 3, 3] = 1
        SE3[:, i] = loc if i == 0 else SE3[:, parent[i]] @ loc
    j = [j_tpose[:, 0]]
    for i in range(1, 16):                                                                         # :285-288
        P = SE3[:, parent[i]]
        j.append(np.einsum('bij,bj->bi', P[:, :3, :3], j_tpose[:, i]) + P[:, :3, 3])
    SE3_v = np.einsum('vj,bjk->bvk', f(m['weights']), SE3.reshape(bs, 16, 16)).reshape(bs, -1, 4, 4)   # :291
    v_out = np.einsum('bvij,bvj->bvi', SE3_v[:, :, :3, :3], v_tpose) + SE3_v[:, :, :3, 3]               # :293-294
    j = np.stack(j + [v_out[:, t] for t in TIPS], 1)[:, NEW_ORDER]                                      # :296-299
    if center_idx is not None:
        c = j[:, center_idx:center_idx + 1]; v_out = v_out - c; j = j - c
    if scale is not None:
        s = f(scale)[:, None, None]; v_out = v_out * s; j = j * s
    if trans is not None:
        t = f(trans)[:, None]; v_out = v_out + t; j = j + t
    if new_skel:                                                                                        # :316-320
        j = j.copy()
        j[:, 5] = (v_out[:, 63] + v_out[:, 144]) / 2; j[:, 9] = (v_out[:, 271] + v_out[:, 220]) / 2
        j[:, 13] = (v_out[:, 148] + v_out[:, 290]) / 2; j[:, 17] = (v_out[:, 770] + v_out[:, 83]) / 2
    return v_out.astype(np.float32), j.astype(np.float32)


def mano_forward_torch(m, root_rotation, pose, shape, trans=None, scale=None, use_pca=True, center_idx=9, new_skel=False, parent=PARENT):
    """Differentiable torch restatement of the same forward (models/manolayer.py:250-322), used to check the fused backward kernel:
    gradients come from torch autograd over these plain ops, exactly how the reference obtains them.  Pinned against the unmodified
    reference's own gradients by tests/golden/mano_grad_synth.pt (oracle/make_golden.py mano_grad)."""
    import torch
    dt = root_rotation.dtype
    f = lambda a: torch.as_tensor(np.asarray(a, np.float64)).to(dt)
    bs = root_rotation.shape[0]
    if use_pca:
        axis = (pose @ f(m['hands_components'])[:pose.shape[1]] + f(m['hands_mean'])).reshape(-1, 3)
        angle = torch.norm(axis, p=2, dim=1, keepdim=True) + 1e-8                 # :37
        u = axis / angle
        L = torch.zeros((axis.shape[0], 3, 3), dtype=dt)
        L[:, 2, 1] = u[:, 0]; L[:, 1, 2] = -u[:, 0]
        L[:, 0, 2] = u[:, 1]; L[:, 2, 0] = -u[:, 1]
        L[:, 1, 0] = u[:, 2]; L[:, 0, 1] = -u[:, 2]
        rot = (torch.eye(3, dtype=dt)[None] + torch.sin(angle)[..., None] * L + (1 - torch.cos(angle))[..., None] * L.bmm(L)).view(bs, 15, 3, 3)
    else:
        rot = pose.reshape(bs, 15, 3, 3)
    jreg = m['J_regressor']
    jreg = f(jreg.todense() if hasattr(jreg, 'todense') else jreg)
    v_shaped = f(m['v_template'])[None] + torch.einsum('vck,bk->bvc', f(m['shapedirs']), shape)
    j_tpose = torch.einsum('jv,bvc->bjc', jreg, v_shaped)
    pose_shape = (rot - torch.eye(3, dtype=dt)).reshape(bs, 135)
    v_tpose = v_shaped + torch.einsum('vck,bk->bvc', f(m['posedirs']), pose_shape)
    eye = torch.eye(3, dtype=dt)[None]
    G = []
    for i in range(16):
        R = root_rotation if i == 0 else rot[:, i - 1]
        t = torch.einsum('bij,bj->bi', eye - R, j_tpose[:, i])
        if i == 0:
            G.append((R, t))
        else:
            PR, Pt = G[parent[i]]
            G.append((PR.bmm(R), torch.einsum('bij,bj->bi', PR, t) + Pt))
    js = [j_tpose[:, 0]]
    for i in range(1, 16):
        PR, Pt = G[parent[i]]
        js.append(torch.einsum('bij,bj->bi', PR, j_tpose[:, i]) + Pt)
    W = f(m['weights'])
    GR = torch.stack([g[0] for g in G], 1); Gt = torch.stack([g[1] for g in G], 1)
    TR = torch.einsum('vj,bjrc->bvrc', W, GR); Tt = torch.einsum('vj,bjr->bvr', W, Gt)
    v_out = torch.einsum('bvrc,bvc->bvr', TR, v_tpose) + Tt
    j = torch.stack(js + [v_out[:, t] for t in TIPS], 1)[:, NEW_ORDER]
    if center_idx is not None:
        c = j[:, center_idx:center_idx + 1]; v_out = v_out - c; j = j - c
    if scale is not None:
        v_out = v_out * scale[:, None, None]; j = j * scale[:, None, None]
    if trans is not None:
        v_out = v_out + trans[:, None]; j = j + trans[:, None]
    if new_skel:
        mids = {5: (63, 144), 9: (271, 220), 13: (148, 290), 17: (770, 83)}
        j = torch.stack([(v_out[:, mids[k][0]] + v_out[:, mids[k][1]]) / 2 if k in mids else j[:, k] for k in range(21)], 1)
    return v_out, j
