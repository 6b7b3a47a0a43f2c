"""Evaluation-metric kernels (csrc/metrics.cu, rih_eval_metrics) against the CPU oracle (oracle/metrics_ref.py, pinned to the reference's own
evaluation loop) and against the golden produced by executing apps/eval_interhand.py."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _regressors():
    from renderih_b200 import assets as A
    from renderih_b200.metrics import Jr
    J16 = {s: torch.from_numpy(np.asarray(A.synthetic_mano(0, s)['J_regressor'].todense(), dtype='float32')) for s in ('left', 'right')}
    return J16, {s: Jr(J16[s]) for s in ('left', 'right')}


def _close(a, b, rtol, what):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert float((a - b).abs().max()) <= rtol * float(b.abs().max()) + 1e-12, (what, float((a - b).abs().max()), float(b.abs().max()))


def test_eval_metrics_match_oracle_and_reference_golden():
    from oracle import fixtures, metrics_ref
    from renderih_b200.metrics import batch_metrics
    J16, jr = _regressors()
    J21 = {s: metrics_ref.joint_regressor21(J16[s]) for s in J16}
    for s in J21:
        assert torch.equal(jr[s].J_regressor.cpu(), J21[s])
    case = fixtures.make_eval_case(6)
    ours = batch_metrics(jr, *[case[k].cuda() for k in ('pred_left', 'pred_right', 'gt_left', 'gt_right')])
    ref = metrics_ref.eval_batch(J21['left'].double(), J21['right'].double(),          # fp64 oracle, intended (quirk-free) Procrustes layout
                                 *[case[k].double() for k in ('pred_left', 'pred_right', 'gt_left', 'gt_right')], batch_quirk=False)
    for side in ('left', 'right'):
        for k in ('orijoint_loss', 'orivert_loss', 'joints_loss', 'verts_loss'):
            _close(ours[k][side], ref[side][k], 2e-6, (side, k))
        for k in ('pajoints_loss', 'paverts_loss'):
            _close(ours[k][side], ref[side][k], 2e-5, (side, k))
    _close(ours['mrrpe'], ref['mrrpe'], 1e-5, 'mrrpe')
    assert torch.isnan(ours['cdev'][-1])
    _close(ours['cdev'][:-1], ref['cdev'][:-1], 2e-6, 'cdev')
    for k in ('double_pa_joint', 'double_pa_mesh', 'double_joint', 'double_mesh'):
        _close(ours[k], ref[k], 2e-5, k)
    # the numbers the unmodified reference produced (first batch of 4: no layout quirk there)
    gold = torch.load(os.path.join(GOLD, 'eval_metrics_synth.pt'), weights_only=False)
    for side in ('left', 'right'):
        for k in ('orijoint_loss', 'orivert_loss', 'joints_loss', 'verts_loss'):
            _close(ours[k][side], gold['per_element'][side][k], 5e-6, ('gold', side, k))
        for k in ('pajoints_loss', 'paverts_loss'):
            _close(ours[k][side][:4], gold['per_element'][side][k][:4], 5e-5, ('gold', side, k))
    _close(ours['mrrpe'], gold['mrrpe'], 1e-5, 'gold mrrpe')
    _close(ours['cdev'][:-1], gold['cdev'][:-1], 5e-6, 'gold cdev')
    _close(ours['double_pa_joint'], gold['double_per_sample']['pa_joint'], 5e-5, 'gold double pa joint')
    _close(ours['double_pa_mesh'], gold['double_per_sample']['pa_mesh'], 5e-5, 'gold double pa mesh')


def test_eval_accumulator_summary_and_invariances():
    """EvalMetrics over uneven batches == oracle summarize; Procrustes errors are invariant to a similarity transform of the prediction and
    never exceed the root-aligned ones by construction of the optimum; per_element=False gives the same per-sample numbers."""
    from oracle import fixtures, metrics_ref
    from renderih_b200.metrics import EvalMetrics, batch_metrics
    J16, jr = _regressors()
    J21 = {s: metrics_ref.joint_regressor21(J16[s]).double() for s in J16}
    case = fixtures.make_eval_case(70, seed=123)
    keys = ('pred_left', 'pred_right', 'gt_left', 'gt_right')
    acc = EvalMetrics(jr)
    refs = []
    for lo, hi in ((0, 64), (64, 69), (69, 70)):
        acc.update(*[case[k][lo:hi].cuda() for k in keys])
        refs.append(metrics_ref.eval_batch(J21['left'], J21['right'], *[case[k][lo:hi].double() for k in keys], batch_quirk=False))
    s, r = acc.summary(), metrics_ref.summarize(refs)
    for side in ('left', 'right'):
        for k in ('ori_mpjpe', 'ori_mpvpe', 'mpjpe', 'mpvpe', 'pa_mpjpe', 'pa_mpvpe'):
            assert abs(s['%s_%s' % (k, side)] - r[side][k]) <= 2e-5 * r[side][k], (side, k, s['%s_%s' % (k, side)], r[side][k])
    for k, rk in (('double_pa_mpjpe', 'double_pa_joint'), ('double_pa_mpvpe', 'double_pa_mesh'), ('double_mpjpe', 'double_joint'), ('double_mpvpe', 'double_mesh'),
                  ('mrrpe', 'mrrpe'), ('cdev', 'cdev')):
        assert abs(s[k] - r[rk]) <= 2e-5 * abs(r[rk]), (k, s[k], r[rk])
    mask = torch.arange(70) % 3 == 0
    sm = acc.summary(mask)
    full = batch_metrics(jr, *[case[k].cuda() for k in keys], per_element=True)
    assert abs(sm['mpjpe_left'] - float(full['joints_loss']['left'][mask.cuda()].mean() * 1000)) <= 1e-4 * sm['mpjpe_left']
    lean = batch_metrics(jr, *[case[k].cuda() for k in keys], per_element=False)
    assert torch.equal(full['sample'], lean['sample']) and 'verts_loss' not in lean
    # similarity invariance of the aligned errors
    g = torch.Generator().manual_seed(3)
    Q, _ = torch.linalg.qr(torch.randn(70, 3, 3, generator=g))
    Q = Q * torch.sign(torch.det(Q))[:, None, None]
    moved = {k: case[k] for k in keys}
    for k in ('pred_left', 'pred_right'):
        moved[k] = 1.37 * case[k].bmm(Q.transpose(1, 2)) + torch.randn(70, 1, 3, generator=g)
    mv = batch_metrics(jr, *[moved[k].cuda() for k in keys], per_element=False)
    for col in (4, 5):
        _close(mv['sample'][:, :, col], full['sample'][:, :, col], 2e-4, 'similarity invariance col %d' % col)
    assert bool((full['sample'][:, :, 4] <= full['sample'][:, :, 0] * (1 + 1e-5)).all()) and bool((full['sample'][:, :, 5] <= full['sample'][:, :, 1] * (1 + 1e-5)).all())


def test_eval_metrics_rejects_cpu_tensors():
    from renderih_b200.metrics import batch_metrics
    _, jr = _regressors()
    z = torch.zeros(1, 778, 3)
    with pytest.raises(RuntimeError):
        batch_metrics(jr, z, z, z, z)


def test_eval_metrics_timing_report(capsys):
    """Not a pass/fail criterion: records the per-batch cost of the two metric kernels next to the oracle's formulation run eagerly on the
    same GPU (the reference's own formulation: torch ops + batched SVD), for DESIGN.md."""
    from oracle import fixtures, metrics_ref
    from renderih_b200.metrics import batch_metrics
    J16, jr = _regressors()
    J21 = {s: metrics_ref.joint_regressor21(J16[s]).cuda() for s in J16}
    case = {k: v.cuda() for k, v in fixtures.make_eval_case(64, seed=7).items()}
    keys = ('pred_left', 'pred_right', 'gt_left', 'gt_right')

    def timed(fn, n):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    t_kernel = timed(lambda: batch_metrics(jr, *[case[k] for k in keys], per_element=True), 50)

    def eager():
        for side, p, g in (('left', 'pred_left', 'gt_left'), ('right', 'pred_right', 'gt_right')):
            metrics_ref.hand_metrics(J21[side], case[p], case[g], batch_quirk=False)
    t_eager = timed(eager, 5)
    with capsys.disabled():
        print('\n[eval metrics, batch 64] fused kernels %.3f ms/batch; eager torch formulation of the single-hand part alone %.2f ms/batch' % (t_kernel, t_eager))
    assert t_kernel > 0
