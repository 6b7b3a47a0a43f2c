"""Run ONLY the bench's roofline kernel (3x3 128->128 conv at 64x64, batch 64, forward) a few times -- for
`ncu --set full -k regex:gemm_tc_persistent -c 1 python tools/roofline_kernel.py [--mode tf32rn]`."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from renderih_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--mode', default='tf32rn')
args = ap.parse_args()
ops.set_gemm_mode(args.mode, args.mode)
N, H, C = 64, 64, 128
x = torch.randn(N * H * H, C, device='cuda')
w = (torch.randn(C, C, 3, 3, device='cuda') * 0.03).contiguous(memory_format=torch.channels_last)
for _ in range(3):
    y = ops.conv2d(x, w, None, N, H, H, stride=1, pad=1)
torch.cuda.synchronize()
print('ok', float(y.abs().mean()))
