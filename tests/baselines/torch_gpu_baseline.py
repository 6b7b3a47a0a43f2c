"""Reference-equivalent eager PyTorch on the GPU (the op sequence of the reference model via oracle/model_ref.py run on
CUDA tensors with autograd): images/s of forward + calc_loss_GCN + backward at batch 64.  This is the 'reference PyTorch-GPU'
number north_star compares against (cuDNN TF32 convolutions by default, exactly like the reference); it is NOT the bench's
reference arm (that one is the CPU path).   usage: python tests/baselines/torch_gpu_baseline.py [--batch 64] [--no-tf32]"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fixtures, model_ref  # noqa: E402
from renderih_b200 import assets as A   # noqa: E402
from renderih_b200.model import load_model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--no-tf32', action='store_true')
    ap.add_argument('--steps', type=int, default=10)
    args = ap.parse_args()
    if args.no_tf32:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    a = A.synthetic_assets(0)
    sd = fixtures.init_state_dict(load_model(assets=a).state_dict())
    sd = {k: v.cuda() for k, v in sd.items()}
    for k, v in sd.items():
        if v.is_floating_point() and 'running_' not in k and k not in ('decoder.dense_coor', 'decoder.unsample_layer.weight'):
            v.requires_grad_(True)
    Ap = model_ref.prepare_assets(a)
    for s in Ap:
        Ap[s]['L'] = [l.cuda() for l in Ap[s]['L']]
    la = fixtures.make_loss_assets(a, A.synthetic_mano(0, 'left'), A.synthetic_mano(0, 'right'))
    la = {s: {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()} for s, d in la.items()}
    B = args.batch
    img = torch.randn(B, 3, 256, 256, device='cuda')
    labels = {k: v.cuda() for k, v in fixtures.make_labels(B).items()}
    params = [v for v in sd.values() if v.requires_grad]

    def step(train=True):
        for p in params:
            p.grad = None
        out = model_ref.model_forward(sd, Ap, img, training=True, dropout=0.05)
        loss = model_ref.calc_loss_GCN(out, labels, la)
        loss.backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    with torch.no_grad():
        for _ in range(3):
            model_ref.model_forward(sd, Ap, img, training=False)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            model_ref.model_forward(sd, Ap, img, training=False)
        e1.record()
        torch.cuda.synchronize()
    msf = e0.elapsed_time(e1) / args.steps
    print('torch eager GPU (tf32 conv=%s) batch %d: fwd+loss+bwd %.2f ms/step = %.1f img/s ; eval fwd %.2f ms = %.1f img/s'
          % (not args.no_tf32, B, ms, B / ms * 1e3, msf, B / msf * 1e3))


if __name__ == '__main__':
    main()
