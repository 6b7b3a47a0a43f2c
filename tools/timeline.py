#!/usr/bin/env python
"""Device timeline of one captured training step (the step bench.py times): per-kernel start / duration / stream from CUPTI through
torch.profiler (nsys is not in this image), written as a compact CSV + a summary of where the wall time goes:
union-busy time, idle gaps, per-stream busy time, concurrency histogram, per-kernel-class exclusive time ("only this class running").

    python tools/timeline.py [--encoder resnet50] [--batch 64] [--out gpurun_out/timeline.csv]

Numbers under a profiler are NOT bench values (CUPTI adds per-kernel overhead); the structure (what overlaps with what, where the GPU idles,
which chain is critical) is what this is for.
"""
import argparse
import csv
import json
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_step(encoder, batch, gemm_mode='ref'):
    import numpy as np
    import torch
    from renderih_b200 import _lib, assets as A, ops
    from renderih_b200.config import load_cfg
    from renderih_b200.loss import GraphLoss, calc_loss_GCN
    from renderih_b200.model import load_model
    from renderih_b200.train import TrainStep
    _lib.load()
    conv_mode, lin_mode = {'ref': ('tf32c', 'tf32x3')}.get(gemm_mode, (gemm_mode, gemm_mode))
    ops.set_gemm_mode(conv_mode, lin_mode)
    cfg = load_cfg()
    a = A.synthetic_assets(0)
    torch.manual_seed(cfg.SEED)
    if encoder in ('graph', 'newgraph'):
        from renderih_b200 import myhand
        build = myhand.load_graph_model if encoder == 'graph' else myhand.load_new_model
        model = build(cfg, assets=a, mano_assets={s: A.synthetic_mano(0, s) for s in ('left', 'right')}).cuda().train()
    else:
        cfg.MODEL.ENCODER_TYPE = encoder
        model = load_model(cfg, assets=a).cuda().train()
    model.decoder.unsample_layer.weight.requires_grad_(False)
    B = batch
    g = torch.Generator().manual_seed(cfg.SEED)
    img = torch.randn(B, 3, 256, 256, generator=g).cuda()
    lab = {k: (torch.randn(*s, generator=g) * 0.05).cuda() for k, s in (('v3d_l', (B, 778, 3)), ('v3d_r', (B, 778, 3)), ('root_rel', (B, 3)))}
    lab.update({k: (torch.rand(B, 778, 2, generator=g) * 256).cuda() for k in ('v2d_l', 'v2d_r')})
    ml, mr = A.synthetic_mano(0, 'left'), A.synthetic_mano(0, 'right')
    jl = torch.from_numpy(np.asarray(ml['J_regressor'].todense(), dtype='float32'))
    jr = torch.from_numpy(np.asarray(mr['J_regressor'].todense(), dtype='float32'))
    gl, gr = GraphLoss(jl, ml['f'], 4, 'cuda'), GraphLoss(jr, mr['f'], 4, 'cuda')
    conv = model.decoder.converter
    z = torch.zeros(B, 21, 3, device='cuda')

    def loss_fn(out):
        return calc_loss_GCN(cfg, 0, gl, gr, conv['left'], conv['right'], out[0], out[1], out[2], out[3], None, None, None,
                             lab['v2d_l'], z[..., :2], lab['v2d_r'], z[..., :2], lab['v3d_l'], z, lab['v3d_r'], z, lab['root_rel'], 256)[0]
    step = TrainStep(model, loss_fn, img, lr=cfg.TRAIN.LR, weight_decay=cfg.TRAIN.weight_decay, labels=lab)
    step.capture(warmup=2)
    return step


def build_forward(encoder, batch, gemm_mode='ref'):
    """Eval-mode forward under torch.no_grad(), captured into a CUDA graph (what `bench.py --config forward` times)."""
    import torch
    from renderih_b200 import _lib, assets as A, ops
    from renderih_b200.config import load_cfg
    from renderih_b200.model import load_model
    _lib.load()
    conv_mode, lin_mode = {'ref': ('tf32c', 'tf32x3')}.get(gemm_mode, (gemm_mode, gemm_mode))
    ops.set_gemm_mode(conv_mode, lin_mode)
    cfg = load_cfg()
    cfg.MODEL.ENCODER_TYPE = encoder
    torch.manual_seed(cfg.SEED)
    model = load_model(cfg, assets=A.synthetic_assets(0)).cuda().eval()
    img = torch.randn(batch, 3, 256, 256, generator=torch.Generator().manual_seed(cfg.SEED)).cuda()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        for _ in range(2):
            model(img)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph), torch.no_grad():
        out = model(img)
    keep = (model, img, out)

    def step():
        keep  # noqa: B018  (static tensors of the graph stay alive)
        graph.replay()
    return step


def klass(name):
    n = re.sub(r'^void ', '', name)
    n = re.sub(r'rih::|tc::|at::native::|<unnamed>::', '', n)
    m = re.match(r'(gemm_tc_persistent_kernel)<(\d+), *(?:\(bool\))?(\w+), *(?:\(bool\))?(\w+), *(\w+)<[^>]*>+, *(\d)>', n)
    if m:
        prod = m.group(5)
        arith = {'1': 'tf32', '2': 'tf32rn', '3': '3xtf32'}[m.group(6)]
        return 'tc:%s:%s' % (prod.replace('Producer', ''), arith)
    return re.sub(r'[<(].*', '', n)[:48]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--encoder', default='resnet50')
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'timeline.csv'))
    ap.add_argument('--forward', action='store_true', help='eval-mode forward (bench.py --config forward) instead of the training step')
    args = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile
    step = build_forward(args.encoder, args.batch) if args.forward else build_step(args.encoder, args.batch)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    tmp = args.out + '.trace.json'
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    prof.export_chrome_trace(tmp)
    with open(tmp) as f:
        tr = json.load(f)
    ks = [e for e in tr['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset') and e.get('ph') == 'X']
    os.remove(tmp)
    t0 = min(e['ts'] for e in ks)
    rows = sorted(((e['ts'] - t0, e['dur'], e.get('args', {}).get('stream', e.get('tid')), e['name'], e.get('cat')) for e in ks), key=lambda r: r[0])
    with open(args.out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['start_us', 'dur_us', 'stream', 'class', 'name'])
        for s, d, st, n, c in rows:
            w.writerow(['%.3f' % s, '%.3f' % d, st, klass(n) if c == 'kernel' else c, n[:160]])
    summarize(rows)


def summarize(rows, top=28):
    evs = [(s, s + d, st, klass(n) if c == 'kernel' else c) for s, d, st, n, c in rows]
    end = max(e[1] for e in evs)
    pts = []
    for i, (s, e, st, k) in enumerate(evs):
        pts.append((s, 1, i)); pts.append((e, -1, i))
    pts.sort()
    active = set()
    last = 0.0
    busy = 0.0
    conc = defaultdict(float)
    excl = defaultdict(float)      # time during which ONLY kernels of this class run
    share = defaultdict(float)     # time apportioned 1/n among concurrently running kernels
    for t, kind, i in pts:
        dt = t - last
        if dt > 0:
            n = len(active)
            conc[n] += dt
            if n:
                busy += dt
                ks = set(evs[j][3] for j in active)
                if len(ks) == 1:
                    excl[next(iter(ks))] += dt
                for j in active:
                    share[evs[j][3]] += dt / n
        last = t
        if kind == 1:
            active.add(i)
        else:
            active.discard(i)
    print('timeline: %d device activities, span %.3f ms, union busy %.3f ms, idle %.3f ms' % (len(evs), end / 1e3, busy / 1e3, (end - busy) / 1e3))
    print('concurrency (kernels in flight -> ms): ' + ', '.join('%d: %.2f' % (k, v / 1e3) for k, v in sorted(conc.items())))
    per_stream = defaultdict(lambda: [0, 0.0])
    for s, e, st, k in evs:
        per_stream[st][0] += 1; per_stream[st][1] += e - s
    print('per stream (id: launches, busy ms): ' + ', '.join('%s: %d, %.2f' % (st, v[0], v[1] / 1e3) for st, v in sorted(per_stream.items(), key=lambda kv: -kv[1][1])))
    tot = defaultdict(lambda: [0, 0.0])
    for s, e, st, k in evs:
        tot[k][0] += 1; tot[k][1] += e - s
    print('%-40s %6s %9s %9s %9s' % ('class', 'count', 'sum ms', 'share ms', 'alone ms'))
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -share[kv[0]])[:top]:
        print('%-40s %6d %9.3f %9.3f %9.3f' % (k, n, us / 1e3, share[k] / 1e3, excl[k] / 1e3))
    # gaps: idle intervals of the whole device
    gaps = []
    cur_end = 0.0
    for s, e, st, k in sorted(evs):
        if s > cur_end:
            gaps.append(s - cur_end)
        cur_end = max(cur_end, e)
    gaps.sort(reverse=True)
    print('idle gaps: %d, total %.3f ms, largest %s us' % (len(gaps), sum(gaps) / 1e3, ['%.1f' % g for g in gaps[:8]]))


if __name__ == '__main__':
    main()
