"""HRNet-w48 encoder variant (BASELINE config 5; reference models/encoder.py:176-352 + models/model_zoo/hrnet.py) on the CUDA path:
kernel-level parity of the HRNet-specific kernels against plain torch fp32, and end-to-end parity of the whole model against the
CPU oracle and the golden vectors the unmodified reference produced (tests/golden/model_hrnet48_synth_b2.pt).

Tolerances: as tests/test_model_gpu.py (eval forward 2e-5 relative to each tensor's max; train-mode forward 5e-3, loss 1e-3, gradient
norms 2e-2 at batch 2 where BatchNorm statistics over 2x8x8 samples make the graph ill-conditioned)."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import fixtures, model_ref
from renderih_b200 import assets as rih_assets

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
FWD_TOL, TRAIN_FWD_TOL, GRAD_TOL = 2e-5, 5e-3, 2e-2
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel_err(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def rows(t):      # NCHW -> NHWC rows
    N, C, H, W = t.shape
    return t.permute(0, 2, 3, 1).reshape(N * H * W, C).contiguous()


def nchw(r, N, H, W):
    return r.view(N, H, W, -1).permute(0, 3, 1, 2)


def flat(out):
    result, params, hlist, other = out
    d = {}
    for side in ('left', 'right'):
        d['verts3d_' + side] = result['verts3d'][side]; d['verts2d_' + side] = result['verts2d'][side]
        d['scale_' + side] = params['scale'][side]; d['trans2d_' + side] = params['trans2d'][side]
        d['v3c_' + side] = hlist[0]['verts3d'][side]; d['v2c_' + side] = hlist[0]['verts2d'][side]
        d['v3list_' + side] = other['verts3d_MANO_list'][side][0]; d['v2list_' + side] = other['verts2d_MANO_list'][side][0]
    for k in ('hms', 'mask', 'dense'):
        t = other[k] if other[k].dim() == 4 else other[k][:, None]
        d[k + '_sub'] = t[:, :, ::8, ::8]; d[k + '_mean'] = t.mean(dim=(2, 3))
        d[k] = other[k]
    return d


# ----------------------------------------------------------------------------- kernels
@pytest.mark.parametrize('N,H0,chans', [(2, 64, (48, 96, 192, 384)), (3, 16, (8, 4, 12, 16)), (1, 8, (4, 4))])
def test_hr_concat_bilinear_matches_torch(N, H0, chans):
    from renderih_b200 import ops
    g = torch.Generator().manual_seed(5)
    xs = [torch.randn(N, c, H0 >> i, H0 >> i, generator=g).cuda().requires_grad_(True) for i, c in enumerate(chans)]
    ref = torch.cat([xs[0]] + [F.interpolate(x, size=(H0, H0), mode='bilinear', align_corners=True) for x in xs[1:]], 1)
    rs = [rows(x.detach()).requires_grad_(True) for x in xs]
    y = ops.hr_concat(rs, N, H0)
    assert rel_err(nchw(y, N, H0, H0), ref) < 1e-6
    w = torch.randn(ref.shape, generator=g).cuda()
    (ref * w).sum().backward()
    (y * rows(w)).sum().backward()
    for x, r in zip(xs, rs):
        assert rel_err(nchw(r.grad, N, x.shape[2], x.shape[3]), x.grad) < 2e-6


@pytest.mark.parametrize('relu', [True, False])
def test_fuse_sum_nearest_upsample_matches_torch(relu):
    from renderih_b200 import ops
    g = torch.Generator().manual_seed(6)
    N, H, C = 2, 16, 24
    ts = [torch.randn(N, C, H // f, H // f, generator=g).cuda().requires_grad_(True) for f in (1, 1, 2, 8)]
    ref = ts[0] + ts[1] + F.interpolate(ts[2], scale_factor=2, mode='nearest') + F.interpolate(ts[3], scale_factor=8, mode='nearest')
    if relu:
        ref = F.relu(ref)
    rs = [rows(t.detach()).requires_grad_(True) for t in ts]
    y = ops.fuse_sum(rs, [1, 1, 2, 8], N, H, relu=relu)
    assert rel_err(nchw(y, N, H, H), ref) == 0.0          # same association order: bit-exact
    w = torch.randn(ref.shape, generator=g).cuda()
    (ref * w).sum().backward()
    (y * rows(w)).sum().backward()
    for t, r in zip(ts, rs):
        assert rel_err(nchw(r.grad, N, t.shape[2], t.shape[3]), t.grad) < 1e-6


def test_biased_conv_batchnorm_block_matches_torch():
    """Conv2d(bias) -> BN -> ReLU of hrnet_mid's downsamp modules (models/encoder.py:304-312), exact-fp32 and tensor-core paths"""
    from renderih_b200 import ops
    from renderih_b200.hrnet import conv_bn, _cbr_seq
    torch.manual_seed(3)
    seq = _cbr_seq(64, 96, 3, 2, bias=True).cuda().train()
    x = torch.randn(4, 64, 16, 16, device='cuda')
    ref = F.relu(F.batch_norm(F.conv2d(x, seq[0].weight, seq[0].bias, stride=2, padding=1), None, None, seq[1].weight, seq[1].bias, True, 0.1, 1e-5))
    for mode, tol in (('simt', 1e-5), ('tf32x3', 1e-4)):
        ops.set_gemm_mode(mode, mode)
        try:
            y, h = conv_bn(rows(x), seq[0], seq[1], 4, 16, True)
        finally:
            ops.set_gemm_mode('simt', 'simt')
        assert h == 8 and rel_err(nchw(y, 4, 8, 8), ref) < tol, mode


# ----------------------------------------------------------------------------- whole model
@pytest.fixture(scope='module')
def gold():
    return torch.load(os.path.join(GOLD, 'model_hrnet48_synth_b2.pt'), weights_only=False)


@pytest.fixture(scope='module')
def setup(gold):
    from renderih_b200.config import load_cfg
    from renderih_b200.model import load_model
    a = rih_assets.synthetic_assets(0)
    cfg = load_cfg(None)
    cfg.MODEL.ENCODER_TYPE = 'hrnet48'
    model = load_model(cfg, assets=a)
    sd = fixtures.init_state_dict(model.state_dict())
    assert fixtures.checksum(sd) == gold['weights_sha256']
    model.load_state_dict(sd)
    return a, sd, model.cuda()


def test_hrnet_forward_eval_matches_oracle_and_reference_golden(gold, setup):
    a, sd, model = setup
    model.eval()
    img = fixtures.make_image(gold['batch'])
    with torch.no_grad():
        out = flat(model(img.cuda()))
        ora = flat(model_ref.model_forward({k: v.clone() for k, v in sd.items()}, model_ref.prepare_assets(a), img, training=False))
    worst = {}
    for k, v in ora.items():
        assert out[k].shape == v.shape, (k, out[k].shape, v.shape)
        worst[k] = rel_err(out[k], v)
    print('hrnet48 eval fwd rel errs vs oracle:', {k: '%.2e' % e for k, e in worst.items()})
    for k, e in worst.items():
        assert e < FWD_TOL, (k, e)
    for k, v in gold['eval'].items():
        assert rel_err(out[k], v) < FWD_TOL, ('golden', k, rel_err(out[k], v))


def test_hrnet_forward_eval_tensor_core_modes(gold, setup):
    """bench arithmetic (tf32c convolutions + 3xTF32 Linears, tolerance 1e-2) and all-3xTF32 (fp32-faithful, tolerance 1e-3)"""
    from renderih_b200 import ops
    a, sd, model = setup
    model.load_state_dict(sd)
    model.eval()
    img = fixtures.make_image(gold['batch'])
    for conv, lin, tol in (('tf32c', 'tf32x3', 1e-2), ('tf32x3', 'tf32x3', 1e-3)):
        ops.set_gemm_mode(conv, lin)
        try:
            with torch.no_grad():
                out = flat(model(img.cuda()))
        finally:
            ops.set_gemm_mode('simt', 'simt')
        errs = {k: rel_err(out[k], v) for k, v in gold['eval'].items()}
        print('hrnet48 %s/%s eval rel errs vs reference golden:' % (conv, lin), {k: '%.2e' % e for k, e in errs.items()})
        for k, e in errs.items():
            assert e < tol, (conv, lin, k, e)


def _oracle_loss_backward(a, sd0, training, dtype):
    """CPU oracle forward + calc_loss_GCN + backward in `dtype`; returns (loss, {key: grad})."""
    sd = {k: (v.clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd0.items()}
    for k, v in sd.items():
        if v.is_floating_point() and 'running_' not in k and k not in ('decoder.dense_coor', 'decoder.unsample_layer.weight'):
            v.requires_grad_(True)
    A = model_ref.prepare_assets(a)
    for side in ('left', 'right'):
        A[side]['L'] = [l.to(dtype) for l in A[side]['L']]
    out = model_ref.model_forward(sd, A, fixtures.make_image(2).to(dtype), training=training, dropout=0.0)
    la = fixtures.make_loss_assets(a, rih_assets.synthetic_mano(0, 'left'), rih_assets.synthetic_mano(0, 'right'))
    for side in la:
        la[side]['J21'] = la[side]['J21'].to(dtype)
    loss = model_ref.calc_loss_GCN(out, {k: v.to(dtype) for k, v in fixtures.make_labels(2).items()}, la)
    loss.backward()
    return float(loss), {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}


def _ours_loss_backward(a, model, training):
    model.train(training)
    for m in model.modules():
        if hasattr(m, 'p'):
            m.p = 0.0       # dropout RNG streams cannot match torch's: parity runs use TRAIN.dropout = 0
    model.decoder.unsample_layer.weight.requires_grad_(False)
    la = fixtures.make_loss_assets(a, rih_assets.synthetic_mano(0, 'left'), rih_assets.synthetic_mano(0, 'right'))
    la_cuda = {s: {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()} for s, d in la.items()}
    model.zero_grad()
    out = model(fixtures.make_image(2).cuda())
    loss = model_ref.calc_loss_GCN(out, {k: v.cuda() for k, v in fixtures.make_labels(2).items()}, la_cuda)   # loss graph in torch (caller side)
    loss.backward()
    return out, loss


ZERO_GRAD = {'mid_model.downsamp_modules.%d.0.bias' % i for i in range(3)} | {'mid_model.final_layer.0.bias'}


def test_hrnet_backward_eval_mode_matches_oracle(gold, setup):
    """Every backward kernel of the HRNet path against the fp32 CPU oracle with BatchNorm in eval mode (running statistics): a
    far better conditioned graph (still ~150 layers deep with gradients of magnitude 1e7), so per-tensor gradient norms are held to
    1e-2 and directions to cos > 0.999 (measured worst: printed)."""
    a, sd, model = setup
    model.load_state_dict(sd)
    _, loss = _ours_loss_backward(a, model, training=False)
    l_ref, g_ref = _oracle_loss_backward(a, sd, False, torch.float32)
    assert abs(float(loss) - l_ref) / l_ref < 1e-4
    params = dict(model.named_parameters())
    worst, worst_cos = (0.0, None), (1.0, None)
    for k, g in g_ref.items():
        if k.endswith('w_ks.bias'):
            continue
        mine = params[k].grad
        assert mine is not None, k
        n = float(g.norm())
        if n < 1e-6:
            continue
        e = abs(float(mine.norm()) - n) / n
        worst = max(worst, (e, k))
        assert e < 1e-2, (k, e)
        cos = float(F.cosine_similarity(mine.detach().cpu().flatten().double(), g.flatten().double(), dim=0))
        worst_cos = min(worst_cos, (cos, k))
        assert cos > 0.999, (k, cos)
    print('hrnet48 eval-mode backward: worst grad-norm rel err %.2e at %s; worst cosine %.6f at %s' % (worst + worst_cos))


def test_hrnet_forward_backward_train_matches_reference_golden(gold, setup):
    """Train mode (batch statistics) at batch 2: BatchNorm over as few as 2x8x8 samples through ~150 layers makes the gradient
    ill-conditioned -- the unmodified reference's own fp32 gradients are 6 % (median) ... 10 % away from the fp64 evaluation of the same
    graph (measured: see DESIGN.md "Numerics").  So forward / loss are held to the reference golden (5e-3 / 1e-3), and each gradient
    norm must be as close to the fp64 truth as the reference's fp32 run is (3x its error + 2e-2)."""
    a, sd, model = setup
    model.load_state_dict(sd)
    out, loss = _ours_loss_backward(a, model, training=True)
    fo = flat(out)
    errs = {k: rel_err(fo[k], v) for k, v in gold['train']['out'].items()}
    print('hrnet48 train fwd rel errs vs reference golden:', {k: '%.2e' % e for k, e in errs.items()})
    for k, e in errs.items():
        assert e < TRAIN_FWD_TOL, ('train fwd', k, e)
    print('loss ours %.6f reference %.6f' % (float(loss), gold['train']['loss']))
    assert abs(float(loss) - gold['train']['loss']) / gold['train']['loss'] < 1e-3
    assert rel_err(model.encoder.hrnet.bn1.running_mean, gold['train']['bn1_running_mean']) < 1e-4
    assert rel_err(model.encoder.hrnet.bn1.running_var, gold['train']['bn1_running_var']) < 1e-4
    params = dict(model.named_parameters())
    for k in gold['train']['no_grad_keys']:
        g = params[k].grad
        assert g is None or float(g.abs().max()) == 0.0, k
    _, g64 = _oracle_loss_backward(a, sd, True, torch.float64)
    ref_err = {k: abs(g['norm'] - float(g64[k].norm())) / max(float(g64[k].norm()), 1e-6) for k, g in gold['train']['grads'].items()}
    med = sorted(ref_err.values())[len(ref_err) // 2]
    worst = (0.0, None)
    for k, g in gold['train']['grads'].items():
        mine = params[k].grad
        assert mine is not None, k
        if k.endswith('w_ks.bias') or k in ZERO_GRAD:
            continue   # mathematically zero gradients (softmax shift invariance / bias in front of a train-mode BatchNorm)
        n64 = float(g64[k].norm())
        e = abs(float(mine.norm()) - n64) / max(n64, 1e-6)
        worst = max(worst, (e, k))
        assert e < 3 * max(ref_err[k], med) + GRAD_TOL, (k, e, ref_err[k], float(mine.norm()), g['norm'], n64)
    print('hrnet48 train: worst grad-norm rel err vs fp64 truth %.2e at %s (reference fp32 median %.2e)' % (worst + (med,)))
