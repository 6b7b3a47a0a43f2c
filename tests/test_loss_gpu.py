"""Fused GraphLoss kernel (csrc/loss.cu, rih_graph_loss_fwd/_bwd) against the CPU oracle's calc_loss_GCN in fp64 (reference core/Loss.py:103-162,
201-277) and against this package's own torch formulation: value, per-term dictionary, gradients of all eight predicted tensors."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(B, seed, spread):
    from renderih_b200 import assets as A
    from renderih_b200.config import load_cfg
    from renderih_b200.loss import GraphLoss
    from renderih_b200.model import GCN_vert_convert
    from oracle import fixtures
    cfg = load_cfg()
    a = A.synthetic_assets(0)
    ml, mr = A.synthetic_mano(0, 'left'), A.synthetic_mano(0, 'right')
    la = fixtures.make_loss_assets(a, ml, mr)
    J = {s: torch.from_numpy(np.asarray(m['J_regressor'].todense(), dtype='float32')) for s, m in (('left', ml), ('right', mr))}
    gl, gr = GraphLoss(J['left'], ml['f'], 4, 'cuda'), GraphLoss(J['right'], mr['f'], 4, 'cuda')
    conv = {s: GCN_vert_convert(778, a[s + '_graph']['graph_perm_reverse'], a[s + '_graph']['graph_perm']) for s in ('left', 'right')}
    g = torch.Generator().manual_seed(seed)
    lab = {'v3d_l': torch.randn(B, 778, 3, generator=g) * 0.1, 'v3d_r': torch.randn(B, 778, 3, generator=g) * 0.1, 'root_rel': torch.randn(B, 3, generator=g) * 0.1,
           'v2d_l': torch.rand(B, 778, 2, generator=g) * 256, 'v2d_r': torch.rand(B, 778, 2, generator=g) * 256}
    pred = {}
    for s in ('l', 'r'):
        pred['v3p_' + s] = lab['v3d_' + s] + torch.randn(B, 778, 3, generator=g) * spread       # spread 1.5: both Smooth-L1 branches are hit
        pred['v2p_' + s] = lab['v2d_' + s] + torch.randn(B, 778, 2, generator=g) * 20
        pred['v3c_' + s] = torch.randn(B, 252, 3, generator=g) * spread
        pred['v2c_' + s] = torch.rand(B, 252, 2, generator=g) * 256
    return cfg, gl, gr, conv, la, lab, pred


def _pack(pred):
    result = {'verts3d': {'left': pred['v3p_l'], 'right': pred['v3p_r']}, 'verts2d': {'left': pred['v2p_l'], 'right': pred['v2p_r']}}
    hlist = [{'verts3d': {'left': pred['v3c_l'], 'right': pred['v3c_r']}, 'verts2d': {'left': pred['v2c_l'], 'right': pred['v2c_r']}}]
    return result, hlist


def _product(cfg, epoch, gl, gr, conv, lab, pred, fused):
    from renderih_b200.loss import calc_loss_GCN
    os.environ['RIH_FUSED_LOSS'] = '1' if fused else '0'
    try:
        p = {k: v.cuda().requires_grad_(True) for k, v in pred.items()}
        L = {k: v.cuda() for k, v in lab.items()}
        result, hlist = _pack(p)
        z = torch.zeros(L['v3d_l'].shape[0], 21, 3, device='cuda')
        total, _, mano, coarse = calc_loss_GCN(cfg, epoch, gl, gr, conv['left'], conv['right'], result, None, hlist, None, None, None, None,
                                               L['v2d_l'], z[..., :2], L['v2d_r'], z[..., :2], L['v3d_l'], z, L['v3d_r'], z, L['root_rel'], 256)
        (total * 1.7).backward()                       # non-unit upstream gradient
        return total.detach().cpu(), {k: v.detach().cpu() for k, v in mano.items()}, {k: [x.detach().cpu() for x in v] for k, v in coarse.items()}, \
            {k: v.grad.cpu() for k, v in p.items()}
    finally:
        os.environ.pop('RIH_FUSED_LOSS', None)


@pytest.mark.parametrize('B,epoch,spread', [(2, 0, 1.5), (5, 60, 0.3), (64, 60, 1.5)])
def test_fused_graph_loss_matches_oracle_and_torch(B, epoch, spread):
    from oracle import model_ref
    cfg, gl, gr, conv, la, lab, pred = _setup(B, 11 + B, spread)
    tot_f, mano_f, coarse_f, g_f = _product(cfg, epoch, gl, gr, conv, lab, pred, True)
    tot_t, mano_t, coarse_t, g_t = _product(cfg, epoch, gl, gr, conv, lab, pred, False)
    # fp64 oracle
    p64 = {k: v.double().requires_grad_(True) for k, v in pred.items()}
    result, hlist = _pack(p64)
    la64 = {s: {'J21': d['J21'].double(), 'faces': d['faces'], 'perm': d['perm']} for s, d in la.items()}
    w = cfg.LOSS_WEIGHT
    tot_o = model_ref.calc_loss_GCN((result, None, hlist, None), {k: v.double() for k, v in lab.items()}, la64, epoch=epoch, w3d=w.DATA.LABEL_3D,
                                    w2d=w.DATA.LABEL_2D, w_norm=w.GRAPH.NORM.NORMAL, w_edge=w.GRAPH.NORM.EDGE, norm_epoch=w.GRAPH.NORM.NORM_EPOCH)
    (tot_o * 1.7).backward()
    tot_o = tot_o.detach()
    assert abs(float(tot_f) - float(tot_o)) <= 2e-5 * abs(float(tot_o)), (float(tot_f), float(tot_o))
    assert abs(float(tot_f) - float(tot_t)) <= 2e-5 * abs(float(tot_t))
    for k in mano_t:
        assert abs(float(mano_f[k]) - float(mano_t[k])) <= 2e-5 * abs(float(mano_t[k])) + 1e-9, k
    for k in coarse_t:
        assert len(coarse_f[k]) == len(coarse_t[k]) == 1
        assert abs(float(coarse_f[k][0]) - float(coarse_t[k][0])) <= 2e-5 * abs(float(coarse_t[k][0])), k
    for k in pred:
        ref = p64[k].grad
        err_f = (g_f[k].double() - ref).norm() / ref.norm()
        err_t = (g_t[k].double() - ref).norm() / ref.norm()
        assert err_f <= max(2e-5, 3 * float(err_t)), (k, float(err_f), float(err_t))
        assert (g_f[k].double() - ref).abs().max() <= 1e-4 * ref.abs().max() + 1e-9, k


def test_fused_graph_loss_in_cuda_graph():
    """The fused loss must be capturable (TrainStep captures forward + loss + backward) and replay on new inputs."""
    cfg, gl, gr, conv, la, lab, pred = _setup(3, 5, 1.0)
    from renderih_b200.loss import calc_loss_GCN
    L = {k: v.cuda() for k, v in lab.items()}
    p = {k: v.cuda().requires_grad_(True) for k, v in pred.items()}
    z = torch.zeros(3, 21, 3, device='cuda')

    def run():
        for v in p.values():
            v.grad = None
        result, hlist = _pack(p)
        t = calc_loss_GCN(cfg, 60, gl, gr, conv['left'], conv['right'], result, None, hlist, None, None, None, None,
                          L['v2d_l'], z[..., :2], L['v2d_r'], z[..., :2], L['v3d_l'], z, L['v3d_r'], z, L['root_rel'], 256)[0]
        t.backward()
        return t.detach(), p['v3p_r'].grad
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        run()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        t_static, g_static = run()
    with torch.no_grad():
        p['v3p_r'].add_(0.25)
    graph.replay()
    torch.cuda.synchronize()
    t_graph, g_graph = t_static.clone(), g_static.clone()
    t_eager, g_eager = run()
    assert torch.allclose(t_graph, t_eager, rtol=1e-6)
    # the face-gradient scatter uses shared-memory atomics: summation order (hence the last bits) varies from launch to launch
    assert float((g_graph - g_eager).abs().max()) <= 1e-5 * float(g_eager.abs().max())


@pytest.mark.parametrize('epoch', [0, 60])
def test_fused_mano_loss_matches_reference_golden(epoch):
    """mano_loss_GCN with the fused mesh terms against the golden produced by the unmodified core/Loss_mano.mano_loss_GCN
    (tests/golden/mano_loss_synth.pt: total, per-term values, gradients of every prediction tensor) and against the torch formulation."""
    from oracle import fixtures
    from renderih_b200 import assets as A
    from renderih_b200.config import load_cfg
    from renderih_b200.loss import ManoLoss, mano_loss_GCN
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'mano_loss_synth.pt'), weights_only=False)[epoch]
    cfg = load_cfg()
    ml, mr = A.synthetic_mano(0, 'left'), A.synthetic_mano(0, 'right')
    J = {s: torch.from_numpy(np.asarray(m['J_regressor'].todense(), dtype='float32')) for s, m in (('left', ml), ('right', mr))}
    gl, gr = ManoLoss(J['left'], ml['f'], 4, 'cuda'), ManoLoss(J['right'], mr['f'], 4, 'cuda')
    res = {}
    for fused in (True, False):
        os.environ['RIH_FUSED_LOSS'] = '1' if fused else '0'
        try:
            pred, lab = fixtures.make_mano_loss_case(2)
            pred = {k: v.cuda().requires_grad_(True) for k, v in pred.items()}
            lab = {k: v.cuda() for k, v in lab.items()}
            result, params, hlist, other = fixtures.mano_loss_inputs(pred)
            z = torch.zeros(2, 21, 3, device='cuda')
            total, _, terms, _ = mano_loss_GCN(cfg, epoch, gl, gr, None, None, result, params, hlist, other, None, None, None,
                                               lab['v2d_l'], z[..., :2], lab['v2d_r'], z[..., :2], lab['v3d_l'], z, lab['v3d_r'], z,
                                               lab['root_rel'], 256, lab['lp_gt'], lab['ls_gt'], lab['rp_gt'], lab['rs_gt'])
            total.backward()
            res[fused] = (float(total.detach()), {k: float(v.detach()) for k, v in terms.items()}, {k: v.grad.cpu() for k, v in pred.items()})
        finally:
            os.environ.pop('RIH_FUSED_LOSS', None)
    for fused in (True, False):
        total, terms, grads = res[fused]
        assert abs(total - gold['total']) <= 2e-5 * abs(gold['total']), (fused, total, gold['total'])
        for k, v in gold['terms'].items():
            assert abs(terms[k] - v) <= 2e-5 * abs(v) + 1e-9, (fused, k)
        for k, g in gold['grads'].items():
            assert float((grads[k] - g).norm()) <= 2e-5 * float(g.norm()) + 1e-12, (fused, k)
