"""One training step (forward + calc_loss_GCN + backward + AdamW) bracketed by cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python tools/profile_step.py --batch 64 --gemm-mode tf32x3
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--gemm-mode', default='simt')
    ap.add_argument('--fwd-only', action='store_true')
    ap.add_argument('--encoder', default='resnet50', choices=['resnet50', 'hrnet48'])
    args = ap.parse_args()
    from renderih_b200 import assets as A, ops
    from renderih_b200.config import load_cfg
    from renderih_b200.loss import GraphLoss, calc_loss_GCN
    from renderih_b200.model import load_model
    from renderih_b200.train import FlatParams, trainable_used_params
    cm, lm = {'ref': ('tf32c', 'tf32x3'), 'refrn': ('tf32rn', 'tf32x3')}.get(args.gemm_mode, (args.gemm_mode, args.gemm_mode))
    ops.set_gemm_mode(cm, lm)
    cfg = load_cfg()
    cfg.MODEL.ENCODER_TYPE = args.encoder
    a = A.synthetic_assets(0)
    torch.manual_seed(88)
    model = load_model(cfg, assets=a).cuda()
    model = model.eval() if args.fwd_only else model.train()       # --fwd-only = the eval forward of bench.py --config forward
    model.decoder.unsample_layer.weight.requires_grad_(False)
    B = args.batch
    img = torch.randn(B, 3, 256, 256, device='cuda')
    lab = {k: torch.randn(B, 778, 3, device='cuda') * 0.05 for k in ('v3d_l', 'v3d_r')}
    lab.update({k: torch.rand(B, 778, 2, device='cuda') * 256 for k in ('v2d_l', 'v2d_r')})
    lab['root_rel'] = torch.randn(B, 3, device='cuda') * 0.05
    ml, mr = A.synthetic_mano(0, 'left'), A.synthetic_mano(0, 'right')
    jl = torch.from_numpy(np.asarray(ml['J_regressor'].todense(), dtype='float32'))
    jr = torch.from_numpy(np.asarray(mr['J_regressor'].todense(), dtype='float32'))
    gl, gr = GraphLoss(jl, ml['f'], 4, 'cuda'), GraphLoss(jr, mr['f'], 4, 'cuda')
    conv = model.decoder.converter
    z = torch.zeros(B, 21, 3, device='cuda')

    def loss_fn(out):
        return calc_loss_GCN(cfg, 0, gl, gr, conv['left'], conv['right'], out[0], out[1], out[2], out[3], None, None, None,
                             lab['v2d_l'], z[..., :2], lab['v2d_r'], z[..., :2], lab['v3d_l'], z, lab['v3d_r'], z, lab['root_rel'], 256)[0]

    groups = [g for m in model.modules() if hasattr(m, 'fused_param_groups') for g in m.fused_param_groups()]      # as train.TrainStep does
    fp = FlatParams(trainable_used_params(model, loss_fn, img), groups=groups)

    def step():
        if args.fwd_only:
            with torch.no_grad():
                model(img)
            return
        fp.zero_grad()
        loss_fn(model(img)).backward()
        fp.adamw_step()

    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == '__main__':
    main()
