#!/usr/bin/env python
"""Benchmark of the RenderIH hot path on B200: images/sec of one training step (forward + calc_loss_GCN + backward +
AdamW, + one NCCL gradient all-reduce when N > 1) at batch 64 per GPU, 256x256 synthetic images (BASELINE.json configs[2]).

    python bench.py --gpus N --steps K --warmup W            # ours (torchrun launches N ranks for N > 1)
    python bench.py --impl reference ...                     # the reference algorithm's CPU port on the host cores

Prints ONE JSON line on rank 0 (contract in the task statement): value = device-timed images/s with inputs resident in
HBM, e2e = same step driven through the public API from pinned HOST buffers (H2D copy of every batch + D2H loss read),
roofline = the dominant kernel timed alone with CUDA events, cpu_baseline = the oracle port on a bounded CPU sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# FlopCounterMode on the reference (algorithmic 2*MAC FLOPs per image, forward+backward / forward)
FLOPS = {'resnet50': (50.450e9, 17.721e9), 'hrnet48': (168.090e9, 56.201e9),
         'graph': (38.422e9, 13.268e9),     # common/myhand graph variant (ResNet-50 trunk), FlopCounterMode on the reference built on CPU
         'newgraph': (38.464e9, 13.283e9)}  # + ParamRegressor / MANO tail (matmul FLOPs only; the ManoLayer is ~1.2 MFLOP per hand)
FLOPS_PER_IMG_FWD_BWD, FLOPS_PER_IMG_FWD = FLOPS['resnet50']


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=None, help='images per GPU (default 64; 32 for --encoder hrnet48 = BASELINE.json configs[4])')
    ap.add_argument('--encoder', default='resnet50', choices=['resnet50', 'hrnet48', 'graph', 'newgraph'],
                    help='resnet50 = BASELINE.json configs[2] (the headline metric), hrnet48 = configs[4] (MODEL.ENCODER_TYPE), '
                         'graph = the common/myhand default model variant (SURVEY 8 f1; ResNet-50 trunk, batch 64), '
                         'newgraph = the same with the ParamRegressor + MANO tail, trained with mano_loss_GCN')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--gemm-mode', default='ref', choices=['ref', 'refrn', 'simt', 'tf32', 'tf32rn', 'tf32c', 'tf32x3'],
                    help="arithmetic of the conv / Linear GEMMs; 'ref' = the reference's own GPU numerics class: TF32 convolutions (as "
                         "cuDNN runs them by default; here truncating TF32 with the 7.05e-4 mean shrinkage compensated in the epilogue, "
                         "measured as accurate as round-to-nearest) + fp32-faithful (3xTF32) nn.Linear GEMMs; 'refrn' = same with "
                         "round-to-nearest TF32 convolutions")
    ap.add_argument('--cpu-batch', type=int, default=4, help='bounded CPU sample size for cpu_baseline / --impl reference')
    ap.add_argument('--skip-cpu-baseline', action='store_true')
    a = ap.parse_args()
    if a.batch is None:
        a.batch = 32 if a.encoder == 'hrnet48' else 64
    return a


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return {'hbm_gbs': p['hbm_gbs'], 'tflops': p['bf16_tflops'], 'tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']),
                'src': 'measured (MEASURED_PEAKS.json)'}
    except Exception:
        return {'hbm_gbs': 6650.0, 'tflops': 1590.0, 'tflops_sustained': 1400.0, 'src': 'fallback (B200_PROFILING.md)'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'samples': len(sm), 'reasons': sorted(reasons)}


def metric_name(batch, world, encoder='resnet50'):
    return 'images/sec fwd+bwd @batch%d 256x256 (training step: fwd + %s + bwd + AdamW%s)' % (
        batch, 'mano_loss_GCN' if encoder == 'newgraph' else 'calc_loss_GCN', ' + NCCL grad all-reduce' if world > 1 else '')


def workload_name(encoder, batch):
    return 'BASELINE.json configs[%d]: HandNET_GCN %s cfg, batch %d/GPU, 256x256, train mode (batch-stat BN, dropout 0.05), random-init weights, ' \
           'synthetic graph/MANO assets' % (4 if encoder == 'hrnet48' else 2, {'graph': 'common/myhand graph variant (ResNet50 trunk)', 'newgraph': 'common/myhand newgraph variant (graph + ParamRegressor + MANO tail, mano_loss_GCN)'}.get(encoder, encoder), batch)


def host_cores():
    """CPU threads this process can actually run on: the scheduler affinity mask, capped by the cgroup CPU quota (a container on
    a 128-thread host may own far fewer; oversubscribing torch's intra-op pool there makes the CPU baseline tens of times slower
    than it really is)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        with open('/sys/fs/cgroup/cpu.max') as f:
            q, per = f.read().split()[:2]
        if q != 'max':
            n = max(1, min(n, int(float(q) / float(per) + 0.999)))
    except Exception:
        pass
    return n


def cpu_port_step_time(batch, steps=1, warmup=0, budget_s=1e9, encoder='resnet50'):
    """The reference algorithm's CPU port (oracle/model_ref.py): forward + calc_loss_GCN + backward, fp32, on the host threads that
    give it the best throughput (a forward-only calibration pass picks among the usable-core count and its halvings down to 16)."""
    import torch
    from oracle import fixtures, model_ref
    from renderih_b200 import assets as A
    from renderih_b200.config import load_cfg
    from renderih_b200.model import load_model
    a = A.synthetic_assets(0)
    cfg = load_cfg()
    if encoder in ('graph', 'newgraph'):
        from renderih_b200 import myhand
        build = myhand.load_graph_model if encoder == 'graph' else myhand.load_new_model
        sd = fixtures.init_state_dict(build(cfg, assets=a, mano_assets={s: A.synthetic_mano(0, s) for s in ('left', 'right')}).state_dict())
    else:
        cfg.MODEL.ENCODER_TYPE = encoder
        sd = fixtures.init_state_dict(load_model(cfg, assets=a).state_dict())
    avail = host_cores()
    cands, c = [], avail
    while c >= 16 and len(cands) < 4:
        cands.append(c); c //= 2
    cands = cands or [avail]
    if len(cands) > 1:
        Ap0, im0, best = model_ref.prepare_assets(a), fixtures.make_image(batch), None
        for c in cands:
            torch.set_num_threads(c)
            t0 = time.perf_counter()
            with torch.no_grad():
                model_ref.model_forward({k: v.clone() for k, v in sd.items()}, Ap0, im0, training=True, dropout=0.0)
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, c)
        cores = best[1]
    else:
        cores = cands[0]
    torch.set_num_threads(cores)
    for k, v in sd.items():
        if v.is_floating_point() and 'running_' not in k and '.mano_' not in k and k not in ('decoder.dense_coor', 'decoder.unsample_layer.weight'):
            v.requires_grad_(True)
    Ap = model_ref.prepare_assets(a)
    if encoder == 'newgraph':
        Ap = fixtures.add_mano_assets(Ap, A.synthetic_mano(0, 'left'), A.synthetic_mano(0, 'right'))
    la = fixtures.make_loss_assets(a, A.synthetic_mano(0, 'left'), A.synthetic_mano(0, 'right'))
    img, labels = fixtures.make_image(batch), fixtures.make_labels(batch)
    opt = torch.optim.AdamW([v for v in sd.values() if v.requires_grad], lr=cfg.TRAIN.LR, weight_decay=cfg.TRAIN.weight_decay)
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        if times and time.perf_counter() - t_begin > budget_s:
            break
        for v in sd.values():
            v.grad = None
        t0 = time.perf_counter()
        out = model_ref.model_forward(sd, Ap, img, training=True, dropout=0.05)
        if encoder == 'newgraph':      # mesh terms on the MANO vertices + pose / shape terms (the loss graph is a few hundred small ops either way)
            loss = model_ref.calc_loss_GCN((out[0], out[1], [], out[3]), labels, la) + sum((d['mano_pose'] ** 2).mean() + (d['mano_shape'] ** 2).mean()
                                                                                              for d in out[3]['verts3d_MANO_list'].values())
        else:
            loss = model_ref.calc_loss_GCN(out, labels, la)
        loss.backward()
        opt.step()
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
    return sum(times) / len(times), cores, len(times)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # bounded: the CPU port needs tens of seconds per step on a big host, so at most ~3 steps / ~150 s are timed
    t, cores, nsteps = cpu_port_step_time(args.cpu_batch, steps=max(1, min(args.steps, 3)), warmup=0, budget_s=150.0, encoder=args.encoder)
    v = args.cpu_batch / t
    line = {'impl': 'reference', 'metric': metric_name(args.batch, args.gpus, args.encoder), 'value': v, 'unit': 'images/s',
            'n_gpus': args.gpus, 'steps': nsteps, 'warmup': 0, 'ms_per_step': t * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(args.encoder, args.batch), 'global_batch': args.batch * args.gpus, 'parallelism': 'dp%d' % args.gpus,
                       'reference_arm': 'the reference algorithm (oracle/model_ref.py port: same torch CPU ops as the reference graph) on the host cores, '
                                        'each step a bounded sample of batch %d images of the batch-%d workload' % (args.cpu_batch, args.batch)},
            'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                             'sample': 'batch %d fwd+calc_loss_GCN+bwd+AdamW, oracle/model_ref.py on torch CPU fp32, %d of %d usable threads' % (args.cpu_batch, cores, host_cores())},
            'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def dominant_kernel_roofline(torch, batch, pk, conv_mode):
    """Time the single most expensive kernel of the step alone: the 3x3 128->128 conv at 64x64 (2.42 GF/img fwd),
    in the arithmetic mode the step runs its convolutions in.  `traffic` = DRAM bytes per launch of that kernel from the
    committed `ncu --set full` capture (profiles/r01_roofline_kernel.json), when present."""
    from renderih_b200 import ops
    N, H, C = batch, 64, 128
    x = torch.randn(N * H * H, C, device='cuda')
    w = (torch.randn(C, C, 3, 3, device='cuda') * 0.03).contiguous(memory_format=torch.channels_last)
    for _ in range(3):
        ops.conv2d(x, w, None, N, H, H, stride=1, pad=1)
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.conv2d(x, w, None, N, H, H, stride=1, pad=1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * N * H * H * C * 9 * C
    ach = flops / (ms * 1e-3) / 1e12
    kname = {'tf32c': 'gemm_tc_persistent_kernel<128,0,0,ConvFwdProducer<128>,1>', 'simt': 'gemm_simt_kernel<128,128,8,8,ConvFwdA,DenseK>', 'tf32': 'gemm_tc_persistent_kernel<128,0,0,ConvFwdProducer<128>,1>',
             'tf32rn': 'gemm_tc_persistent_kernel<128,0,0,ConvFwdProducer<128>,2>', 'tf32x3': 'gemm_tc_persistent_kernel<128,0,0,ConvFwdProducer<128>,3>'}[conv_mode]
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'r01_roofline_kernel.json')) as f:
            traffic = json.load(f).get(conv_mode, {}).get('dram_bytes_per_launch')
    except Exception:
        pass
    return {'bound': 'tensor', 'kernel': 'conv3x3 128->128 @64x64 batch %d fwd (%s)' % (batch, kname), 'achieved': ach, 'peak': pk['tflops'],
            'unit': 'TFLOP/s', 'frac': ach / pk['tflops'], 'traffic': traffic, 'ms_per_launch': ms,
            'algorithmic_flops_per_launch': flops, 'algorithmic_bytes_per_launch': 4.0 * (2 * N * H * H * C + 9 * C * C),
            'peak_source': pk['src'] + ' bf16 dense burst (TF32 tensor peak is half of it)'}


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device: there is no CPU fallback for the product path'
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from renderih_b200 import _lib, assets as A
    from renderih_b200.config import load_cfg
    from renderih_b200.loss import GraphLoss, calc_loss_GCN
    from renderih_b200.model import load_model
    from renderih_b200.train import TrainStep
    _lib.load()
    from renderih_b200 import ops as _ops
    conv_mode, lin_mode = {'ref': ('tf32c', 'tf32x3'), 'refrn': ('tf32rn', 'tf32x3')}.get(args.gemm_mode, (args.gemm_mode, args.gemm_mode))
    _ops.set_gemm_mode(conv_mode, lin_mode)
    cfg = load_cfg()
    flops_fb = FLOPS[args.encoder][0]
    a = A.synthetic_assets(0)
    torch.manual_seed(cfg.SEED)
    if args.encoder in ('graph', 'newgraph'):
        from renderih_b200 import myhand
        build = myhand.load_graph_model if args.encoder == 'graph' else myhand.load_new_model
        model = build(cfg, assets=a, mano_assets={s: A.synthetic_mano(0, s) for s in ('left', 'right')}).cuda().train()
    else:
        cfg.MODEL.ENCODER_TYPE = args.encoder
        model = load_model(cfg, assets=a).cuda().train()          # train mode: batch-stat BN, dropout 0.05 (reference defaults)
    model.decoder.unsample_layer.weight.requires_grad_(False)   # MODEL.freeze_upsample
    B = args.batch
    g = torch.Generator().manual_seed(cfg.SEED + rank)
    host_imgs = [torch.randn(B, 3, 256, 256, generator=g).pin_memory() for _ in range(2)]
    lab = {k: (torch.randn(*s, generator=g) * 0.05).cuda() for k, s in (('v3d_l', (B, 778, 3)), ('v3d_r', (B, 778, 3)), ('root_rel', (B, 3)))}
    lab.update({k: (torch.rand(B, 778, 2, generator=g) * 256).cuda() for k in ('v2d_l', 'v2d_r')})
    ml, mr = A.synthetic_mano(0, 'left'), A.synthetic_mano(0, 'right')
    jl = torch.from_numpy(__import__('numpy').asarray(ml['J_regressor'].todense(), dtype='float32'))
    jr = torch.from_numpy(__import__('numpy').asarray(mr['J_regressor'].todense(), dtype='float32'))
    gl, gr = GraphLoss(jl, ml['f'], 4, 'cuda'), GraphLoss(jr, mr['f'], 4, 'cuda')
    conv = model.decoder.converter
    z = torch.zeros(B, 21, 3, device='cuda')
    if args.encoder == 'newgraph':
        from renderih_b200.loss import ManoLoss, mano_loss_GCN
        gl, gr = ManoLoss(jl, ml['f'], 4, 'cuda'), ManoLoss(jr, mr['f'], 4, 'cuda')
        lab.update({k: (torch.randn(B, n, generator=g) * 0.3).cuda() for k, n in (('lp_gt', 48), ('rp_gt', 48), ('ls_gt', 10), ('rs_gt', 10))})

    def loss_fn(out):
        if args.encoder == 'newgraph':
            return mano_loss_GCN(cfg, 0, gl, gr, conv['left'], conv['right'], out[0], out[1], out[2], out[3], None, None, None,
                                 lab['v2d_l'], z[..., :2], lab['v2d_r'], z[..., :2], lab['v3d_l'], z, lab['v3d_r'], z, lab['root_rel'], 256,
                                 lab['lp_gt'], lab['ls_gt'], lab['rp_gt'], lab['rs_gt'])[0]
        return calc_loss_GCN(cfg, 0, gl, gr, conv['left'], conv['right'], out[0], out[1], out[2], out[3], None, None, None,
                             lab['v2d_l'], z[..., :2], lab['v2d_r'], z[..., :2], lab['v3d_l'], z, lab['v3d_r'], z, lab['root_rel'], 256)[0]

    step = TrainStep(model, loss_fn, host_imgs[0].cuda(), lr=cfg.TRAIN.LR, weight_decay=cfg.TRAIN.weight_decay, use_graph=not args.no_graph)
    c0 = _lib.CALLS[0]
    step._eager()
    torch.cuda.synchronize()
    calls_per_step = _lib.CALLS[0] - c0
    if not args.no_graph:
        step.capture(warmup=2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device='cuda')
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    for i in range(max(3, args.warmup)):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev = timed(lambda i: step(), args.steps)                       # inputs already resident in HBM
    losses = []

    step.prefetch(host_imgs[0])

    def e2e_step(i):
        loss = step()                                                    # consumes the batch whose H2D copy was started one step ahead ...
        step.prefetch(host_imgs[(i + 1) % len(host_imgs)])               # ... and starts the next one (pinned memory, copy stream): one H2D per step
        losses.append(float(loss))                                       # D2H read of the step's result (synchronises every step)

    for i in range(2):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    roof = dominant_kernel_roofline(torch, B, pk, conv_mode)
    total_imgs = B * world
    value = total_imgs / (ms_dev * 1e-3)
    e2e = total_imgs / (ms_e2e * 1e-3)
    line = {'metric': metric_name(B, world, args.encoder),
            'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': ms_dev,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': {'simt': 'f32', 'refrn': 'f32 storage; tcgen05 TF32(rn) convolutions (cuDNN-default class of the reference) + 3xTF32 fp32-faithful Linear GEMMs, fp32 accumulate',
                      'tf32': 'tf32 (truncating) conv+Linear, fp32 accumulate/storage', 'tf32c': 'tf32 (truncating, mean-compensated) conv+Linear, fp32 accumulate/storage',
                      'ref': 'f32 storage; tcgen05 TF32 convolutions (truncating + mean-compensated: the accuracy class of the reference\'s cuDNN-TF32 default, measured) + 3xTF32 fp32-faithful Linear GEMMs, fp32 accumulate', 'tf32rn': 'tf32 (rn) conv+Linear, fp32 accumulate/storage',
                      'tf32x3': '3xTF32 (fp32-faithful) conv+Linear, fp32 accumulate/storage'}[args.gemm_mode], 'data': 'synthetic',
            'config': {'workload': workload_name(args.encoder, B),
                       'global_batch': total_imgs, 'parallelism': 'dp%d' % world, 'cuda_graph': not args.no_graph,
                       'l2': 'per-step working set (activations, several GB) >> 126 MB L2; no explicit flush needed',
                       'algorithmic_gflop_per_image': flops_fb / 1e9},
            'achieved_tflops': value * flops_fb / 1e12,
            'e2e': {'value': e2e, 'unit': 'images/s', 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': B * 3 * 256 * 256 * 4, 'd2h_bytes_per_step': 4,
                    'note': 'TrainStep public API from pinned float32 host batches; the H2D copy of step i+1 is issued on a copy stream while step i runs (input double buffering)'},
            'gpu_launches': calls_per_step * args.steps, 'launches_per_step': calls_per_step,
            'clocks': clocks, 'roofline': roof, 'last_loss': losses[-1] if losses else None}
    if not args.skip_cpu_baseline and world == 1:
        t, cores, _ = cpu_port_step_time(args.cpu_batch, steps=1, warmup=0, encoder=args.encoder)
        line['cpu_baseline'] = {'value': args.cpu_batch / t, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                                'sample': 'batch %d fwd+calc_loss_GCN+bwd+AdamW once, oracle/model_ref.py (torch CPU fp32, %d of %d usable threads)' % (args.cpu_batch, cores, host_cores())}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
