"""CPU restatement of the reference loader's per-sample image path (core/loader.py:105-181, utils/manoutils.py:140-261) -- TEST
INFRASTRUCTURE ONLY (never imported by the product package).

    warpAffine (cv2, INTER_LINEAR, BORDER_CONSTANT 0)  ->  brightness noise a*img + b, clip, uint8  ->  horizontal flip
    ->  BGR2RGB, / 255, CHW, Normalize(mean, std)

`cv2.warpAffine` is a third-party dependency of the reference (requirements.txt pins opencv_python 4.6.0.66 / 4.7.0.72; this container
has 4.13.0).  Its published algorithm for 8-bit images, restated in `warp_affine_u8`: the 2x3 matrix is inverted in double precision;
per destination pixel the source position is evaluated in fixed point with 10 fractional bits (AB_BITS) and rounded to 1/32 pixel
(INTER_BITS = 5); the four neighbours are blended with integer weights (32-fy)(32-fx)... that sum to 1024 (OpenCV stores them x32 as
15-bit shorts: every weight is an exact multiple of 32, so `(sum + 2^14) >> 15` equals `(sum/32 + 512) >> 10`), taps outside the image
contribute the border value 0.  Pinned bit-exactly against cv2 itself in tests/test_oracle_golden.py (when cv2 is importable) and against
tests/golden/augment_synth.pt, which oracle/make_golden.py produces with the reference's own `imgUtils` functions + cv2.
"""
import math

import numpy as np

AB_BITS, INTER_BITS = 10, 5
MEAN = np.array([0.485, 0.456, 0.406], np.float32)     # core/loader.py:49-50
STD = np.array([0.229, 0.224, 0.225], np.float32)


def get_affine_mat(theta=0.0, scale=1.0, u=0, v=0, height=256, width=256):
    """imgUtils.get_affine_mat, utils/manoutils.py:183-195 (+ get_rotation_mat :158-170, get_scale_mat :144-155): float32 3x3,
    trans(u, v) @ scale-about-centre @ rotation-about-centre; note the reference's pi = 3.14159."""
    center = np.array([width / 2, height / 2, 1], dtype='float32')
    t = theta * (3.14159 / 180)
    rot = np.zeros((3, 3), dtype='float32')
    rot[0, 0], rot[0, 1], rot[1, 0], rot[1, 1], rot[2, 2] = math.cos(t), -math.sin(t), math.sin(t), math.cos(t), 1.0
    tr = np.matmul((np.identity(3, dtype='float32') - rot), center)
    rot[0, 2], rot[1, 2] = tr[0], tr[1]
    sc = np.zeros((3, 3), dtype='float32')
    sc[0, 0], sc[1, 1], sc[2, 2] = scale, scale, 1.0
    ts = np.matmul((np.identity(3, dtype='float32') - sc), center)
    sc[0, 2], sc[1, 2] = ts[0], ts[1]
    trans = np.identity(3, dtype='float32')
    trans[0, 2], trans[1, 2] = u, v
    return np.matmul(trans, np.matmul(sc, rot))


def rotation_mat3d(theta):
    """imgUtils.get_rotation_mat3d, utils/manoutils.py:172-181"""
    t = theta * (3.14159 / 180)
    R = np.zeros((3, 3), dtype='float32')
    R[0, 0], R[0, 1], R[1, 0], R[1, 1], R[2, 2] = math.cos(t), -math.sin(t), math.sin(t), math.cos(t), 1.0
    return R


def invert_affine(M):
    """cv::warpAffine's in-place inversion of the forward matrix (imgwarp.cpp, `if (!(flags & WARP_INVERSE_MAP))`), double precision."""
    M = np.asarray(M, dtype=np.float64).reshape(2, 3).copy()
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    M[0, 0] = A11; M[0, 1] *= -D; M[1, 0] *= -D; M[1, 1] = A22
    b1 = -M[0, 0] * M[0, 2] - M[0, 1] * M[1, 2]
    b2 = -M[1, 0] * M[0, 2] - M[1, 1] * M[1, 2]
    M[0, 2], M[1, 2] = b1, b2
    return M


def warp_affine_u8(img, M, inverse=None):
    """cv2.warpAffine(img, M, (W, H)) for uint8 HxWxC, default flags (INTER_LINEAR, BORDER_CONSTANT, borderValue 0)."""
    H, W = img.shape[:2]
    Mi = invert_affine(M) if inverse is None else np.asarray(inverse, np.float64).reshape(2, 3)
    scale = float(1 << AB_BITS)
    x = np.arange(W, dtype=np.float64)
    y = np.arange(H, dtype=np.float64)
    rnd = lambda v: np.clip(np.rint(v), -2147483648, 2147483647).astype(np.int64)      # saturate_cast<int>(double) = round half to even
    adelta, bdelta = rnd(Mi[0, 0] * x * scale), rnd(Mi[1, 0] * x * scale)
    round_delta = (1 << AB_BITS) // (1 << INTER_BITS) // 2
    X0 = rnd((Mi[0, 1] * y + Mi[0, 2]) * scale) + round_delta
    Y0 = rnd((Mi[1, 1] * y + Mi[1, 2]) * scale) + round_delta
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    # OpenCV stores the integer part as short (saturate_cast<short>): positions beyond +-32767 are outside any image anyway
    sx, sy = np.clip(X >> INTER_BITS, -32768, 32767), np.clip(Y >> INTER_BITS, -32768, 32767)
    fx, fy = X & 31, Y & 31
    src = img.astype(np.int64)
    acc = np.zeros(img.shape, np.int64)
    for dy, dx, w in ((0, 0, (32 - fy) * (32 - fx)), (0, 1, (32 - fy) * fx), (1, 0, fy * (32 - fx)), (1, 1, fy * fx)):
        yy, xx = sy + dy, sx + dx
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        tap = src[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
        acc += np.where(ok[..., None], tap, 0) * w[..., None]
    return ((acc + 512) >> 10).astype(np.uint8)


def add_noise_u8(img_u8, a, b, scale=255.0):
    """imgUtils.add_noise with noise = 0 (the loader's default, core/loader.py:36) followed by `.astype(np.uint8)` (core/loader.py:142-146):
    a [3] float64 per-channel gains (np.random.uniform), b float64 offset; float64 arithmetic, clip to [0, 255], truncate."""
    out = np.asarray(a, np.float64) * img_u8.astype(np.float32) + float(b)
    return np.clip(out, 0, scale).astype(np.uint8)


def to_network_input(img_u8_bgr, flip=False):
    """core/loader.py:151-152 (cv.flip(img, 1)) and :176-181: BGR->RGB, float32 / 255, CHW, Normalize.  Also returns `ori_img` (:176-177)."""
    img = img_u8_bgr[:, ::-1] if flip else img_u8_bgr
    ori = (img.astype(np.float32) / np.float32(255)).transpose(2, 0, 1)
    rgb = img[..., ::-1].astype(np.float32) / np.float32(255)
    out = (rgb.transpose(2, 0, 1) - MEAN[:, None, None]) / STD[:, None, None]
    return ori, out.astype(np.float32)


def augment_labels(theta, affine, label2d_list, label3d_list):
    """imgUtils.data_augmentation, label part (utils/manoutils.py:233-246): 2-D labels through the affine map, 3-D labels rotated about z."""
    R = rotation_mat3d(theta)
    l2 = [np.matmul(l, affine[0:2, 0:2].T) + affine[0:2, 2:3].T for l in label2d_list]
    l3 = [np.matmul(l, R.T) for l in label3d_list]
    return l2, l3


def process_image(img_u8_bgr, theta, scale, u, v, a=None, b=0.0, flip=False):
    """The image half of handDataset.process_data in train mode (core/loader.py:122-181) for given augmentation draws."""
    M = get_affine_mat(theta, scale, u, v, height=img_u8_bgr.shape[0], width=img_u8_bgr.shape[0])
    img = warp_affine_u8(img_u8_bgr, M[0:2, :])
    if a is not None:
        img = add_noise_u8(img, a, b)
    ori, net = to_network_input(img, flip)
    return img, ori, net, M


def process_labels(hand_dict, theta, affine, flip, img_size=256, bone_length=0.095):
    """The label half of handDataset.process_data (core/loader.py:106-113, 122-128, 183-211): augmentation, root-relative 3-D labels
    (root = joint 9), bone-length normalisation (mean of both hands' |j9 - j0| scaled to `bone_length`), flip (x mirrored, hands swapped).
    Returns {v2d_l, j2d_l, v2d_r, j2d_r, v3d_l, j3d_l, v3d_r, j3d_r, root_rel} as float32 arrays."""
    l2 = [hand_dict['left']['verts2d'], hand_dict['left']['joints2d'], hand_dict['right']['verts2d'], hand_dict['right']['joints2d']]
    l3 = [hand_dict['left']['verts3d'], hand_dict['left']['joints3d'], hand_dict['right']['verts3d'], hand_dict['right']['joints3d']]
    l2, l3 = augment_labels(theta, affine, l2, l3)
    root_left, root_right = l3[1][9], l3[3][9]
    root_rel = root_right - root_left
    l3 = [l3[0] - root_left, l3[1] - root_left, l3[2] - root_right, l3[3] - root_right]
    if bone_length is not None:
        length = (np.linalg.norm(l3[1][9] - l3[1][0]) + np.linalg.norm(l3[3][9] - l3[3][0])) / 2
        s = bone_length / length
        root_rel = root_rel * s
        l3 = [l * s for l in l3]
    root_rel = np.asarray(root_rel, np.float32).copy()
    l2 = [np.asarray(l, np.float32).copy() for l in l2]
    l3 = [np.asarray(l, np.float32).copy() for l in l3]
    if flip:
        root_rel[1:] = -root_rel[1:]
        for i in range(4):
            l2[i][:, 0] = img_size - l2[i][:, 0]
            l3[i][:, 0] = -l3[i][:, 0]
        v2d_r, j2d_r, v2d_l, j2d_l = l2
        v3d_r, j3d_r, v3d_l, j3d_l = l3
    else:
        v2d_l, j2d_l, v2d_r, j2d_r = l2
        v3d_l, j3d_l, v3d_r, j3d_r = l3
    return {'v2d_l': v2d_l, 'j2d_l': j2d_l, 'v2d_r': v2d_r, 'j2d_r': j2d_r, 'v3d_l': v3d_l, 'j3d_l': j3d_l, 'v3d_r': v3d_r, 'j3d_r': j3d_r, 'root_rel': root_rel}
