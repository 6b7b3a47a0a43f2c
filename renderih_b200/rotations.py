"""Rotation / rigid-transform helpers that sit beside the ManoLayer in the reference (models/manolayer.py:20-98, 163-248).

They are host-side conveniences of the layer's API surface (datasets and evaluation code call them on a handful of
rows) -- small batched torch expressions on whatever device the argument lives on.  Each function states the
reference lines whose RESULT it reproduces; the formulations are this package's own (closed-form Rodrigues through
the outer product, level-wise frame propagation instead of a 15-step loop, ...), and `tests/test_host_logic.py`
pins every one of them to values produced by the unmodified reference (tests/golden/mano_helpers_synth.pt).
"""
import torch

# the reference folds obtuse angles with a truncated pi (models/manolayer.py:206-207); kept so that results agree digit for digit
_PI_REF = 3.14159
_EPS = 1e-8


def _unit(v, dim=-1):
    return v / torch.linalg.vector_norm(v, dim=dim, keepdim=True)


def _cross_matrix(k):
    """[k]x for k [N,3] -> [N,3,3]: column c is k x e_c."""
    eye = torch.eye(3, dtype=k.dtype, device=k.device)
    return torch.stack([torch.linalg.cross(k, eye[c].expand_as(k)) for c in range(3)], dim=2)


def rodrigues_batch(axis):
    """Axis-angle [N,3] -> rotation matrices [N,3,3] (result of models/manolayer.py:32-48).
    theta = |a| + 1e-8 is added AFTER the norm (the reference's convention; the loss code adds it before), k = a / theta, and
    R = I + sin(theta) [k]x + (1 - cos(theta)) (k k^T - |k|^2 I)   -- [k]x^2 written through the outer product."""
    theta = torch.linalg.vector_norm(axis, dim=1, keepdim=True) + _EPS
    k = axis / theta
    s, c = torch.sin(theta)[:, :, None], torch.cos(theta)[:, :, None]
    eye = torch.eye(3, dtype=axis.dtype, device=axis.device)
    kk = k[:, :, None] * k[:, None, :] - (k * k).sum(1)[:, None, None] * eye
    return eye + s * _cross_matrix(k) + (1 - c) * kk


def rotmat_to_axis(R):
    """Rotation matrices [..,3,3] -> axis-angle [N,3] (result of ManoLayer.Rmat2axis, models/manolayer.py:186-213).
    The antisymmetric part gives sin(theta) * axis; the angle is asin of its length (clamped to +-(1 - 1e-7)), reflected to
    pi - angle when the cosine recovered from the trace is negative.  The cosine is normalised by |k|^2 - 3 exactly like the
    reference, so the degenerate sin ~ 0 band resolves the same way."""
    R = R.reshape(-1, 3, 3)
    A = 0.5 * (R - R.transpose(1, 2))
    v = torch.stack([A[:, 2, 1], A[:, 0, 2], A[:, 1, 0]], dim=1)
    s = torch.linalg.vector_norm(v, dim=1)
    k = v / (s[:, None] + _EPS)
    cos = 1 - (torch.diagonal(R, dim1=1, dim2=2).sum(1) - 3.0) / ((k * k).sum(1) - 3.0 + _EPS)
    sc = s.clamp(-1 + 1e-7, 1 - 1e-7)
    theta = torch.asin(sc)
    theta = torch.where((cos < 0) & (sc > 0), _PI_REF - theta, theta)
    return theta[:, None] * k


def se3_from(R, t):
    """[N,3,3], [N,3,1] -> homogeneous [N,4,4] (ManoLayer.buildSE3_batch, models/manolayer.py:229-238)."""
    T = R.new_zeros((R.shape[0], 4, 4))
    T[:, :3, :3] = R
    T[:, :3, 3:] = t
    T[:, 3, 3] = 1
    return T


def se3_apply(T, v):
    """Apply homogeneous transforms [N,4,4] to points [N,3] (ManoLayer.SE3_apply, models/manolayer.py:240-248)."""
    return torch.baddbmm(T[:, :3, 3:], T[:, :3, :3], v[:, :, None])[:, :, 0]


def vec2mat(vec):
    """6-D rotation representation [N,6] -> [N,3,3] by Gram-Schmidt, columns (x, y, x cross y) (models/manolayer.py:20-29; the 1e-8 is
    added to each norm)."""
    x = vec[:, 0:3] / (torch.linalg.vector_norm(vec[:, 0:3], dim=1, keepdim=True) + _EPS)
    y = vec[:, 3:6] - (x * vec[:, 3:6]).sum(1, keepdim=True) * x
    y = y / (torch.linalg.vector_norm(y, dim=1, keepdim=True) + _EPS)
    return torch.stack([x, y, torch.linalg.cross(x, y)], dim=2)


def swing_between(z_from, z_to):
    """Rotation [..,3,3] about the common perpendicular that carries unit vector z_from onto z_to (get_trans, models/manolayer.py:51-60):
    with u = unit(z_from x z_to) both frames (u, z x u, z) are orthonormal, and the rotation is new_frame old_frame^T."""
    u = _unit(torch.linalg.cross(z_from, z_to))
    old = torch.stack([u, torch.linalg.cross(z_from, u), z_from], dim=-1)
    new = torch.stack([u, torch.linalg.cross(z_to, u), z_to], dim=-1)
    return new @ old.transpose(-1, -2)


# MANO joint bookkeeping (models/manolayer.py:65-69): child of joint 1..15, the five palm (finger-root) joints in adjacency order
# thumb-side -> little finger, and the permutation that undoes ManoLayer.new_order
_CHILD = (2, 3, 17, 5, 6, 18, 8, 9, 20, 11, 12, 19, 14, 15, 16)
_PALM = (13, 1, 4, 10, 7)
_UNDO_NEW_ORDER = (0, 5, 6, 7, 9, 10, 11, 17, 18, 19, 13, 14, 15, 1, 2, 3, 4, 8, 12, 16, 20)


def build_mano_frame(skel21):
    """Zero-pose local joint frames [bs,15,3,3] (columns = splay, bend, twist axes) from a 21-joint skeleton given in ManoLayer's output
    order (result of build_mano_frame, models/manolayer.py:63-98).
      * twist axis z_i = unit(child_i - joint_i) for the 15 articulated joints;
      * at the five finger roots the splay axis starts from the palm normal -- the normalised sum of the normals of the (up to two)
        palm triangles (wrist, root_k, root_k+1) that touch the finger -- and is re-orthogonalised: y = unit(z cross x), x = y cross z;
      * every other joint inherits its parent's frame, swung by the rotation that carries the parent's twist axis onto its own.
    The reference walks the 15 joints in a loop; the kinematic tree has depth 3, so the propagation is done level by level here."""
    skel = skel21[:, list(_UNDO_NEW_ORDER)]
    bs = skel.shape[0]
    z = _unit(skel[:, list(_CHILD)] - skel[:, 1:16], dim=2)                    # joints 1..15
    z = torch.cat([torch.zeros_like(z[:, :1]), z], dim=1)                      # index by joint id (0 = wrist, unused)
    palm = list(_PALM)
    spokes = skel[:, palm] - skel[:, 0:1]                                       # wrist -> finger roots, [bs,5,3]
    normals = _unit(torch.linalg.cross(spokes[:, :-1], spokes[:, 1:], dim=2), dim=2)   # 4 palm triangles
    acc = torch.zeros((bs, 5, 3), dtype=skel.dtype, device=skel.device)
    acc[:, :-1] += normals
    acc[:, 1:] += normals
    x_p = _unit(acc, dim=2)
    y_p = _unit(torch.linalg.cross(z[:, palm], x_p, dim=2), dim=2)
    x_p = torch.linalg.cross(y_p, z[:, palm], dim=2)
    frames = torch.zeros((bs, 16, 3, 3), dtype=skel.dtype, device=skel.device)
    frames[:, palm] = torch.stack([x_p, y_p, z[:, palm]], dim=3)
    level = palm
    for _ in range(2):                                                          # middle joints, then distal joints
        nxt = [j + 1 for j in level]
        frames[:, nxt] = swing_between(z[:, level], z[:, nxt]) @ frames[:, level]
        level = nxt
    return frames[:, 1:]
