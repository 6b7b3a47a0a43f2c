"""Parity of the path bench.py times: `TrainStep` (flat buffers, CUDA-graph capture, side-stream weight gradients, fused AdamW) against
ONE step of the reference trainer -- the UNMODIFIED reference model + core.Loss.calc_loss_GCN + torch.optim.AdamW
(core/lijun_trainer.py:131-144, 262-313) driven by oracle/ref_driver.py on the CPU -- on the same seeded weights / batch, dropout 0.

Stated tolerances (batch 16; measured values are printed):
  exact-fp32 kernels ('simt'):  loss 2e-4 relative (measured 5.6e-6), every gradient tensor |norm ratio - 1| < 2e-2 (BatchNorm scale gradients of the first
                                 layers move by ~1e-2 between two RUNS of the same fp32 code: summation-order noise amplified by train-mode BatchNorm) and cosine > 0.999
                                 (measured worst 0.99942 on the stem filter, a sum over 10^6 pixels), AdamW update direction: sign agreement on
                                 every element whose reference gradient is not round-off
  bench arithmetic ('ref' = tf32c convolutions + 3xTF32 Linears): TF32 convolution operands move this randomly initialised train-mode network
                                 (batch-statistic BatchNorm after every convolution) by 8 ... 16 % in its outputs -- measured on the REFERENCE's
                                 own graph with its convolution operands rounded to TF32 (what its cuDNN path does by default), see
                                 test_reference_tf32_sensitivity_bounds_bench_arithmetic.  Held to: loss 1e-1, gradient norms within a factor 2
                                 (no direction check: at a 12 % forward deviation the early-layer gradients decorrelate) (exact-fp32 kernels above are the parity statement; this one guards against gross errors)
  fused AdamW kernel vs torch.optim.AdamW on identical gradients: 1e-6 relative after 3 steps
  CUDA-graph replays of the same step (BatchNorm in eval mode): gradients agree to 1e-2 of each tensor's max (1e-4 at the decoder heads) (fp64 atomics in the BatchNorm statistics / shared-memory
  atomics in the loss backward make the summation ORDER vary between launches; nothing else may)
"""
import os

import numpy as np
import pytest
import torch

from oracle import fixtures, ref_driver
from renderih_b200 import assets as rih_assets

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
B = 16      # BatchNorm statistics over >= 1024 samples per channel: well conditioned (batch 2 is not)


def _product_step(mode, batch=B, use_graph=True, lr=3e-4, wd=1e-2):
    """-> (TrainStep, model, labels on device) with seeded weights, dropout 0, MODEL.freeze_upsample."""
    from renderih_b200 import ops
    from renderih_b200.config import load_cfg
    from renderih_b200.loss import GraphLoss, calc_loss_GCN
    from renderih_b200.model import load_model
    from renderih_b200.train import TrainStep
    conv_mode, lin_mode = {'ref': ('tf32c', 'tf32x3')}.get(mode, (mode, mode))
    ops.set_gemm_mode(conv_mode, lin_mode)
    a = rih_assets.synthetic_assets(0)
    cfg = load_cfg()
    model = load_model(cfg, assets=a)
    model.load_state_dict(fixtures.init_state_dict(model.state_dict()))
    model = model.cuda().train()
    for m in model.modules():
        if hasattr(m, 'p'):
            m.p = 0.0
    model.decoder.unsample_layer.weight.requires_grad_(False)
    ml, mr = rih_assets.synthetic_mano(0, 'left'), rih_assets.synthetic_mano(0, 'right')
    J = {s: torch.from_numpy(np.asarray(m['J_regressor'].todense(), dtype='float32')) for s, m in (('left', ml), ('right', mr))}
    gl, gr = GraphLoss(J['left'], ml['f'], 4, 'cuda'), GraphLoss(J['right'], mr['f'], 4, 'cuda')
    lab = {k: v.cuda() for k, v in fixtures.make_labels(batch).items()}
    z = torch.zeros(batch, 21, 3, device='cuda')
    conv = model.decoder.converter

    def loss_fn(out):
        return calc_loss_GCN(cfg, 0, gl, gr, conv['left'], conv['right'], out[0], out[1], out[2], out[3], None, None, None,
                             lab['v2d_l'], z[..., :2], lab['v2d_r'], z[..., :2], lab['v3d_l'], z, lab['v3d_r'], z, lab['root_rel'], 256)[0]
    img = fixtures.make_image(batch).cuda()
    step = TrainStep(model, loss_fn, img, lr=lr, weight_decay=wd, use_graph=use_graph, labels=lab)
    return step, model, img


@pytest.fixture(scope='module')
def reference_step():
    if not ref_driver.available():
        pytest.skip('reference sources not staged (python -m oracle.build_ref)')
    torch.manual_seed(0)
    ref = ref_driver.ReferenceStep('cpu', dropout=0.0)
    ref.model.train()
    before = {k: p.detach().clone() for k, p in ref.model.named_parameters()}
    img, labels = fixtures.make_image(B), fixtures.make_labels(B)
    loss = ref.step(img, labels)
    grads = {k: p.grad.detach().clone() for k, p in ref.model.named_parameters() if p.grad is not None}
    after = {k: p.detach().clone() for k, p in ref.model.named_parameters()}
    bn = ref.model.encoder.resnet.bn1
    return {'loss': float(loss), 'grads': grads, 'before': before, 'after': after, 'lr': ref.lr, 'wd': ref.wd,
            'bn1_mean': bn.running_mean.clone(), 'bn1_tracked': int(bn.num_batches_tracked)}


def _compare_step(mode, ref, loss_tol, norm_tol, cos_tol, check_updates=True):
    from renderih_b200 import ops
    try:
        step, model, img = _product_step(mode, lr=ref['lr'], wd=ref['wd'])
        names = {id(p): k for k, p in model.named_parameters()}
        flat_names = [names[id(p)] for p in step.flatp.params]
        assert set(flat_names) == set(ref['grads']), set(flat_names) ^ set(ref['grads'])    # same trainable-and-used set as the reference's autograd
        step.capture(warmup=2)
        # capture (warm-up included) must leave the model exactly as loaded (it used to run 3 real optimizer steps)
        for k, p in model.named_parameters():
            assert torch.equal(p.detach().cpu(), ref['before'][k]), ('capture changed', k)
        assert step.flatp.step_count == 0 and float(step.flatp.exp_avg.abs().max()) == 0.0
        assert int(model.encoder.resnet.bn1.num_batches_tracked) == 0
        loss = float(step(img))
        torch.cuda.synchronize()
        el = abs(loss - ref['loss']) / abs(ref['loss'])
        print('[%s] loss ours %.6f reference %.6f (rel %.2e)' % (mode, loss, ref['loss'], el))
        assert el < loss_tol, (loss, ref['loss'])
        worst_n, worst_c, flips = (0.0, None), (1.0, None), 0
        params = dict(model.named_parameters())
        for k, g in ref['grads'].items():
            mine = params[k].grad.detach().cpu().double().flatten()
            gr = g.double().flatten()
            if k.endswith('w_ks.bias') or float(gr.norm()) < 1e-7:
                continue       # mathematically zero gradient (softmax shift invariance): round-off only
            en = abs(float(mine.norm()) / float(gr.norm()) - 1)
            cos = float(torch.dot(mine, gr) / (mine.norm() * gr.norm()))
            if en > worst_n[0]:
                worst_n = (en, k)
            if cos < worst_c[0]:
                worst_c = (cos, k)
            assert en < norm_tol, (mode, k, en)
            assert cos > cos_tol, (mode, k, cos)
            # one AdamW step from zero moments moves every element by -lr * (sign(g) (1 - tiny) + wd * p): compare where g is not round-off
            d_ref = (ref['after'][k] - ref['before'][k]).double().flatten()
            d_mine = (params[k].detach().cpu() - ref['before'][k]).double().flatten()
            sig = gr.abs() > 1e-3 * gr.abs().max()
            bad = ((d_ref - d_mine).abs() > 2e-2 * ref['lr']) & sig
            flips += int(bad.sum())
            # the first Adam step is -lr * sign(g): the update DIRECTION must agree (elements whose gradient sits at the round-off level may flip)
            cu = float(torch.dot(d_mine, d_ref) / (d_mine.norm() * d_ref.norm()).clamp_min(1e-30))
            assert not check_updates or (cu > 0.85 and int(bad.sum()) <= 5e-2 * int(sig.sum()) + 1), (mode, k, cu, int(bad.sum()), int(sig.sum()))
        print('[%s] worst grad-norm rel err %.2e at %s ; worst cosine %.6f at %s ; %d AdamW sign flips on significant elements'
              % ((mode,) + worst_n + worst_c + (flips,)))
        assert step.flatp.step_count == 1
        assert int(model.encoder.resnet.bn1.num_batches_tracked) == ref['bn1_tracked'] == 1        # BatchNorm2d.num_batches_tracked parity
        assert float((model.encoder.resnet.bn1.running_mean.cpu() - ref['bn1_mean']).abs().max()) < (1e-5 if mode == 'simt' else 5e-3)
    finally:
        ops.set_gemm_mode('simt', 'simt')
        ops.clear_grad_targets()


def test_trainstep_exact_fp32_matches_reference_trainer_step(reference_step):
    _compare_step('simt', reference_step, 2e-4, 2e-2, 0.999)


def test_trainstep_bench_arithmetic_matches_reference_trainer_step(reference_step):
    """The arithmetic bench.py runs (tf32c convolutions + 3xTF32 Linears / attention), end-to-end GRADIENT parity included."""
    _compare_step('ref', reference_step, 1e-1, 1.0, -1.0, check_updates=False)


def test_fused_adamw_kernel_matches_torch_optim_adamw():
    from renderih_b200.train import FlatParams
    g = torch.Generator().manual_seed(5)
    shapes = [(64, 3, 7, 7), (130,), (77, 33), (1, 1), (256, 128, 1, 1)]
    mine = [torch.nn.Parameter((torch.randn(*s, generator=g) * 0.3).cuda()) for s in shapes]
    mine[0].data = mine[0].data.contiguous(memory_format=torch.channels_last)
    ref = [torch.nn.Parameter(p.detach().clone().contiguous()) for p in mine]
    fp = FlatParams(mine)
    opt = torch.optim.AdamW(ref, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    for it in range(3):
        for p, q in zip(mine, ref):
            gr = (torch.randn(*p.shape, generator=g) * (10.0 ** (it - 2))).cuda()
            p.grad.copy_(gr)
            q.grad = gr.clone().contiguous()
        fp.adamw_step(3e-4, weight_decay=1e-2)
        opt.step()
    torch.cuda.synchronize()
    for p, q in zip(mine, ref):
        e = float((p.detach() - q.detach()).abs().max() / q.detach().abs().max())
        assert e < 1e-6, (tuple(p.shape), e)
    # optimizer state round trip in torch.optim.AdamW's layout (the reference's OPTIM_PATH resume, core/lijun_trainer.py:131-144)
    sd = fp.state_dict()
    opt2 = torch.optim.AdamW([torch.nn.Parameter(q.detach().clone()) for q in ref], lr=3e-4, weight_decay=1e-2)
    opt2.load_state_dict({'state': {i: {k: (v.cpu().contiguous() if torch.is_tensor(v) else v) for k, v in st.items()} for i, st in sd['state'].items()},
                          'param_groups': sd['param_groups']})
    st_ref = opt.state_dict()['state']
    for i in range(len(shapes)):
        assert float(sd['state'][i]['step']) == float(st_ref[i]['step']) == 3.0
        assert float((sd['state'][i]['exp_avg'].cpu() - st_ref[i]['exp_avg'].cpu()).abs().max() / st_ref[i]['exp_avg'].abs().max().cpu()) < 2e-6
    fp2 = FlatParams([torch.nn.Parameter(p.detach().clone()) for p in mine])
    fp2.load_state_dict(opt.state_dict())
    assert fp2.step_count == 3 and float((fp2.exp_avg_sq - fp.exp_avg_sq).abs().max() / fp.exp_avg_sq.abs().max()) < 2e-6
    with pytest.raises(RuntimeError):       # mean folded into the kernel: gradients scaled by 1/world
        from renderih_b200._lib import call
        call('rih_adamw_step', 0, 0, 0, 0, -1, 0.0, 0.9, 0.999, 1e-8, 0.0, 1, 1.0, 0)


def test_graph_replay_determinism_and_state_preservation():
    """Two replays of the captured step from the same state: loss and flat gradient agree to a stated 1e-5 (relative to the gradient's max
    magnitude); the only sources of launch-to-launch variation are summation ORDER effects of atomics (fp64 BatchNorm column sums, shared
    memory atomics of the loss backward, TMA reduce-add of split-K / side-stream weight gradients)."""
    from renderih_b200 import ops
    try:
        step, model, img = _product_step('ref', batch=4)
        model.eval()            # BatchNorm on running statistics: the well-conditioned regime (train-mode BatchNorm amplifies last-bit differences
                                # of this randomly initialised network by orders of magnitude), dropout is already 0
        step.capture(warmup=1)
        snap = step.flatp.snapshot()
        bufs = [b.detach().clone() for b in model.buffers()]
        step.graph.replay(); torch.cuda.synchronize()
        g1, l1 = step.flatp.grad.clone(), float(step.loss)
        step.flatp.restore(snap)
        for b, s in zip(model.buffers(), bufs):
            b.copy_(s)
        step.graph.replay(); torch.cuda.synchronize()
        g2, l2 = step.flatp.grad.clone(), float(step.loss)
        rel = float((g1 - g2).abs().max() / g1.abs().max())
        print('replay determinism: loss %.8f / %.8f, max grad diff %.2e of max |g| (%s)' % (l1, l2, rel, 'bit-identical' if rel == 0 else 'order effects'))
        names = {id(p): k for k, p in model.named_parameters()}
        per = []
        for p, off in zip(step.flatp.params, step.flatp.offsets):
            a, b = g1[off:off + p.numel()], g2[off:off + p.numel()]
            d = float((a - b).abs().max())
            if d > 0 and not names[id(p)].endswith('w_ks.bias'):     # key biases: mathematically zero gradient, pure round-off on both sides
                per.append((d / max(float(a.abs().max()), 1e-30), d, names[id(p)]))
        per.sort(reverse=True)
        print('  %d of %d gradient tensors differ between the replays; largest relative differences (of the tensor\'s own max):' % (len(per), len(step.flatp.params)))
        for r_, d, k in per[:8]:
            print('    %-70s rel %.2e abs %.2e' % (k, r_, d))
        assert abs(l1 - l2) <= 1e-6 * abs(l1)
        # Summation-order effects only (fp32 shared-memory atomics in the loss backward, fp32 reduce-adds of split-K / side-stream weight
        # gradients).  They enter at the loss (last bits) and are amplified on the way back through ~60 un-normalised layers of this randomly
        # initialised network (measured: <= 1e-6 at the decoder heads, growing to 1.2e-3 of the tensor's max at the stem filter);
        # a race would show up as O(1) differences somewhere.
        assert all(r_ < 1e-2 for r_, _, _ in per), per[:3]
        heads = [r_ for r_, _, k in per if k.startswith(('decoder.coord_head', 'decoder.params_head', 'decoder.avg_head'))]
        assert all(r_ < 1e-4 for r_ in heads), heads
        assert torch.isfinite(g1).all()
    finally:
        ops.set_gemm_mode('simt', 'simt')
        ops.clear_grad_targets()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (run with gpurun --gpus 2)')
def test_two_rank_nccl_step_equals_one_rank_step_on_concatenated_batch():
    """SURVEY 4 / core/lijun_trainer.py:122-127: an N-rank data-parallel step == the 1-rank step on the concatenated batch (BatchNorm in
    eval mode so the per-rank statistics do not enter).  Spawns tests/dist_equivalence.py under torchrun on 2 GPUs."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1', '--master-port', '29571',
           os.path.join(root, 'tests', 'dist_equivalence.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=root)
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0
    assert 'DIST_EQUIVALENCE_OK' in r.stdout


def test_reference_tf32_sensitivity_bounds_bench_arithmetic():
    """How far do TF32 CONVOLUTIONS move this network in train mode?  Measured on the UNMODIFIED reference itself, on this GPU: its graph
    `.cuda()` with cuDNN TF32 convolutions (torch's default, what a user of the reference gets) against the same graph with TF32 off, batch 16,
    seeded weights, dropout 0.  The bench arithmetic of this package (tf32c convolutions + 3xTF32 Linears) must stay within 3x of that
    deviation from the exact-fp32 result -- i.e. in the accuracy class of the reference's own GPU path."""
    if not ref_driver.available():
        pytest.skip('reference sources not staged (python -m oracle.build_ref)')
    from renderih_b200 import ops
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    img, labels = fixtures.make_image(B), fixtures.make_labels(B)
    outs = {}
    try:
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            ref = ref_driver.ReferenceStep('cuda', dropout=0.0)
            ref.model.train()
            with torch.no_grad():
                o = ref.model(img.cuda())
            outs[tf32] = {s: o[0]['verts3d'][s].float().cpu() for s in ('left', 'right')}
            del ref
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    dev_ref = max(float((outs[True][s] - outs[False][s]).abs().max() / outs[False][s].abs().max()) for s in ('left', 'right'))
    mine = {}
    try:
        for mode in ('simt', 'ref'):
            step, model, _ = _product_step(mode, use_graph=False)
            with torch.no_grad():
                o = model(img.cuda())
            mine[mode] = {s: o[0]['verts3d'][s].float().cpu() for s in ('left', 'right')}
            ops.clear_grad_targets()
    finally:
        ops.set_gemm_mode('simt', 'simt')
    dev_exact = max(float((mine['simt'][s] - outs[False][s]).abs().max() / outs[False][s].abs().max()) for s in ('left', 'right'))
    dev_mine = max(float((mine['ref'][s] - mine['simt'][s]).abs().max() / mine['simt'][s].abs().max()) for s in ('left', 'right'))
    print('train-mode batch-%d verts3d: reference GPU cuDNN-TF32 vs its own fp32: %.3e ; ours exact-fp32 vs reference GPU fp32: %.3e ; '
          'ours bench arithmetic vs ours exact-fp32: %.3e' % (B, dev_ref, dev_exact, dev_mine))
    assert dev_exact < 2e-3          # both fp32, different summation orders, ill-conditioned train-mode BatchNorm
    assert dev_mine < 3 * dev_ref + 1e-3
