"""GPU input pipeline: uint8 HWC BGR frames -> the normalised float32 NCHW batch `HandNET_GCN.forward` takes.

Replaces the per-sample host work of the reference's loader (`core/loader.py:151-152` cv.flip; `:178-181` cv.cvtColor(BGR2RGB), / 255,
permute(2,0,1), `transforms.Normalize(mean=[0.485,0.456,0.406], std=[0.229,0.224,0.225])`, `:49-50`) with one kernel; bit-identical to
those ops, and the host -> device copy is 4x smaller (uint8).

`augment_u8` adds the training-time augmentation in front of it in the same kernel (`core/loader.py:122-150`, `utils/manoutils.py:183-261`):
cv2.warpAffine (bilinear, constant border) bit-exact with OpenCV's 8-bit fixed-point path, the brightness noise `a * img + b` of
`imgUtils.add_noise`, the flip; `get_affine_mat` / `augment_labels` mirror the reference's host helpers for the matrices and the labels.
"""
import math

import numpy as np
import ctypes

import torch

from ._lib import call

IMAGENET_MEAN = (0.485, 0.456, 0.406)    # core/loader.py:49-50
IMAGENET_STD = (0.229, 0.224, 0.225)


def preprocess_u8(frames, flip=None, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """frames: uint8 [B,H,W,3] BGR on the GPU; flip: optional bool/uint8 [B] (horizontal flip per sample) -> float32 [B,3,H,W]."""
    if not (frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3):
        raise RuntimeError('renderih_b200.preprocess_u8: expected a CUDA uint8 [B,H,W,3] tensor (there is no CPU fallback)')
    frames = frames.contiguous()
    B, H, W, _ = frames.shape
    out = torch.empty((B, 3, H, W), device=frames.device, dtype=torch.float32)
    fl = None
    if flip is not None:
        fl = flip.to(device=frames.device, dtype=torch.uint8).contiguous()
        assert fl.numel() == B
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    call('rih_preprocess_u8', frames.data_ptr(), None if fl is None else fl.data_ptr(), out.data_ptr(), B, H, W, m, s,
         torch.cuda.current_stream(frames.device).cuda_stream)
    return out


def get_affine_mat(theta=0.0, scale=1.0, u=0, v=0, height=256, width=256):
    """imgUtils.get_affine_mat (utils/manoutils.py:183-195): float32 3x3 = translate(u, v) @ scale about the centre @ rotate about the centre
    (theta in degrees, with the reference's pi = 3.14159).  Host arithmetic, a few flops per sample."""
    center = np.array([width / 2, height / 2, 1], dtype='float32')
    t = theta * (3.14159 / 180)
    rot = np.zeros((3, 3), dtype='float32')
    rot[0, 0], rot[0, 1], rot[1, 0], rot[1, 1], rot[2, 2] = math.cos(t), -math.sin(t), math.sin(t), math.cos(t), 1.0
    rot[0:2, 2] = np.matmul(np.identity(3, dtype='float32') - rot, center)[0:2]
    sc = np.zeros((3, 3), dtype='float32')
    sc[0, 0], sc[1, 1], sc[2, 2] = scale, scale, 1.0
    sc[0:2, 2] = np.matmul(np.identity(3, dtype='float32') - sc, center)[0:2]
    trans = np.identity(3, dtype='float32')
    trans[0, 2], trans[1, 2] = u, v
    return np.matmul(trans, np.matmul(sc, rot))


def _invert_affine(M):
    """The inversion cv::warpAffine applies to a forward 2x3 matrix (double precision)."""
    M = np.asarray(M, dtype=np.float64).reshape(2, 3).copy()
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    M[0, 0] = A11; M[0, 1] *= -D; M[1, 0] *= -D; M[1, 1] = A22
    b1 = -M[0, 0] * M[0, 2] - M[0, 1] * M[1, 2]
    b2 = -M[1, 0] * M[0, 2] - M[1, 1] * M[1, 2]
    M[0, 2], M[1, 2] = b1, b2
    return M


def augment_u8(frames, affine, gain=None, offset=None, flip=None, mean=IMAGENET_MEAN, std=IMAGENET_STD, return_ori=False, return_u8=False):
    """frames: uint8 [B,H,W,3] BGR on the GPU; affine: [B,2,3] or [B,3,3] forward matrices as passed to cv.warpAffine (host array);
    gain [B,3] / offset [B]: the `a`, `b` draws of imgUtils.add_noise (float64, host) or None; flip: [B] bool or None.
    -> float32 [B,3,H,W] network input (+ `ori_img` float32 BGR/255 CHW, + the augmented uint8 frames, when asked)."""
    if not (frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3):
        raise RuntimeError('renderih_b200.augment_u8: expected a CUDA uint8 [B,H,W,3] tensor (there is no CPU fallback)')
    frames = frames.contiguous()
    B, H, W, _ = frames.shape
    A = np.asarray(affine, dtype=np.float32).reshape(B, -1, 3)[:, 0:2, :]
    minv = torch.from_numpy(np.stack([_invert_affine(a) for a in A]).reshape(B, 6)).to(frames.device)
    go = None
    if gain is not None:
        go = np.concatenate([np.asarray(gain, np.float64).reshape(B, 3), np.asarray(offset, np.float64).reshape(B, 1)], 1)
        go = torch.from_numpy(go).to(frames.device)
    fl = None if flip is None else torch.as_tensor(flip).to(device=frames.device, dtype=torch.uint8).contiguous()
    out = torch.empty((B, 3, H, W), device=frames.device, dtype=torch.float32)
    ori = torch.empty_like(out) if return_ori else None
    u8 = torch.empty_like(frames) if return_u8 else None
    call('rih_augment_u8', frames.data_ptr(), minv.data_ptr(), None if go is None else go.data_ptr(), None if fl is None else fl.data_ptr(),
         out.data_ptr(), None if ori is None else ori.data_ptr(), None if u8 is None else u8.data_ptr(), B, H, W,
         (ctypes.c_float * 3)(*mean), (ctypes.c_float * 3)(*std), torch.cuda.current_stream(frames.device).cuda_stream)
    res = (out,) + ((ori,) if return_ori else ()) + ((u8,) if return_u8 else ())
    return res[0] if len(res) == 1 else res


def augment_labels(theta, affine, label2d, label3d):
    """imgUtils.data_augmentation, label part (utils/manoutils.py:233-246), batched on the labels' device: label2d [B,N,2] through the
    affine map, label3d [B,N,3] rotated about z by theta (degrees, [B]).  float32 like the reference."""
    A = torch.as_tensor(np.asarray(affine, dtype=np.float32).reshape(len(affine), -1, 3)[:, 0:2, :], device=label2d.device)
    t = torch.as_tensor(np.asarray(theta, dtype=np.float64) * (3.14159 / 180))
    R = torch.zeros(len(affine), 3, 3, dtype=torch.float32)
    R[:, 0, 0], R[:, 0, 1], R[:, 1, 0], R[:, 1, 1], R[:, 2, 2] = torch.cos(t).float(), -torch.sin(t).float(), torch.sin(t).float(), torch.cos(t).float(), 1.0
    return label2d @ A[:, :, 0:2].transpose(1, 2) + A[:, :, 2].unsqueeze(1), label3d @ R.to(label3d.device).transpose(1, 2)


def prepare_labels(hand_dict, theta, affine, flip, img_size=256, bone_length=0.095):
    """The label half of handDataset.process_data (core/loader.py:106-113, 183-211), batched on the labels' device.
    hand_dict: {'left' | 'right': {'verts2d' [B,778,2], 'joints2d' [B,21,2], 'verts3d' [B,778,3], 'joints3d' [B,21,3]}} float32;
    theta [B] degrees, affine [B,3,3] (get_affine_mat), flip [B] bool.  Steps: label augmentation, 3-D labels relative to joint 9 of their
    hand, both hands rescaled so that the mean |j9 - j0| equals `bone_length` (BONE_LENGTH = 0.095, dataset/dataset_utils.py:9), flip = mirror
    x and swap the hands.  -> dict v2d_l, j2d_l, v2d_r, j2d_r, v3d_l, j3d_l, v3d_r, j3d_r, root_rel (the loader's return order)."""
    names2, names3 = ('verts2d', 'joints2d'), ('verts3d', 'joints3d')
    dev = hand_dict['left']['verts3d'].device
    out2, out3 = {}, {}
    for side in ('left', 'right'):
        for n2, n3 in zip(names2, names3):
            out2[side, n2], out3[side, n3] = augment_labels(theta, affine, hand_dict[side][n2].float(), hand_dict[side][n3].float())
    root = {s: out3[s, 'joints3d'][:, 9:10] for s in ('left', 'right')}
    root_rel = (root['right'] - root['left'])[:, 0]
    for s in ('left', 'right'):
        for n in names3:
            out3[s, n] = out3[s, n] - root[s]
    if bone_length is not None:
        length = sum(torch.linalg.norm(out3[s, 'joints3d'][:, 9] - out3[s, 'joints3d'][:, 0], dim=-1) for s in ('left', 'right')) / 2
        sc = (bone_length / length)[:, None]
        root_rel = root_rel * sc
        for k in out3:
            out3[k] = out3[k] * sc[:, :, None]
    f = torch.as_tensor(flip, device=dev, dtype=torch.bool)
    mirror3 = torch.tensor([-1.0, 1.0, 1.0], device=dev)
    root_rel = torch.where(f[:, None], root_rel * torch.tensor([1.0, -1.0, -1.0], device=dev), root_rel)
    res = {'root_rel': root_rel}
    for short, n2, n3 in (('v', 'verts2d', 'verts3d'), ('j', 'joints2d', 'joints3d')):
        l2 = {s: torch.where(f[:, None, None], torch.stack([img_size - out2[s, n2][..., 0], out2[s, n2][..., 1]], -1), out2[s, n2]) for s in ('left', 'right')}
        l3 = {s: torch.where(f[:, None, None], out3[s, n3] * mirror3, out3[s, n3]) for s in ('left', 'right')}
        res[short + '2d_l'] = torch.where(f[:, None, None], l2['right'], l2['left'])
        res[short + '2d_r'] = torch.where(f[:, None, None], l2['left'], l2['right'])
        res[short + '3d_l'] = torch.where(f[:, None, None], l3['right'], l3['left'])
        res[short + '3d_r'] = torch.where(f[:, None, None], l3['left'], l3['right'])
    return res
