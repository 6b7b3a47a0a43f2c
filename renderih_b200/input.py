"""GPU input pipeline: uint8 HWC BGR frames -> the normalised float32 NCHW batch `HandNET_GCN.forward` takes.

Replaces the per-sample host work of the reference's loader (`core/loader.py:151-152` cv.flip; `:178-181` cv.cvtColor(BGR2RGB), / 255,
permute(2,0,1), `transforms.Normalize(mean=[0.485,0.456,0.406], std=[0.229,0.224,0.225])`, `:49-50`) with one kernel; bit-identical to
those ops, and the host -> device copy is 4x smaller (uint8).  The geometric augmentation (cv.warpAffine) stays on the host side.
"""
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
