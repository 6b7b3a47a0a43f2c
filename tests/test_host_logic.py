"""Host-side logic that needs no GPU: module structure / state_dict contract, config, asset generators, fail-loud
behaviour, flat parameter buffers and the 2-rank gradient exchange (gloo)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from renderih_b200 import assets as A

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def model():
    from renderih_b200.model import load_model
    return load_model(assets=A.synthetic_assets(0))


def test_state_dict_contract(model):
    sd = model.state_dict()
    assert len(sd) == 1093                                   # SURVEY.md section 5 (checkpoint / resume)
    assert sum(p.numel() for p in model.parameters()) == 39037091
    gold = torch.load(os.path.join(GOLD, 'model_synth_b2.pt'), weights_only=False)
    names = dict(model.named_parameters())
    for k in list(gold['train']['grads']) + gold['train']['no_grad_keys']:   # every reference parameter name exists here
        assert k in names, k
    assert sd['decoder.dense_coor'].shape == (778, 3) and sd['decoder.unsample_layer.weight'].shape == (778, 252)
    assert 'decoder.dual_gcn.layers.0.graph_left.GCN_blocks.0.graph_L' not in sd      # persistent=False in the reference
    for m in model.modules():
        if isinstance(m, torch.nn.Conv2d):
            assert m.weight.permute(0, 2, 3, 1).is_contiguous()       # kernels read [Cout,R,S,Cin]
    sd2 = {('module.' + k): v for k, v in sd.items()}                  # DDP-prefixed checkpoints (eval_interhand.py:241-250)
    model.load_state_dict({k[7:]: v for k, v in sd2.items()})
    assert model.encoder.resnet.layer1[0].conv2.weight.permute(0, 2, 3, 1).is_contiguous()


def test_reference_attribute_surface(model):
    d = model.decoder
    assert d.get_upsample_weight().shape == (778, 252)
    x = torch.arange(2 * 778 * 3, dtype=torch.float32).view(2, 778, 3)
    g = d.converter['left'].vert_to_GCN(x)
    assert g.shape == (2, 1008, 3)
    assert torch.equal(d.converter['left'].GCN_to_vert(g), x)
    assert d.vNum_in == 63 and d.vNum_out == 252 and d.vNum_all == 1008
    assert hasattr(model, 'encoder') and hasattr(model, 'mid_model')


def test_no_cpu_fallback(model):
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 3, 256, 256))
    from renderih_b200 import ops
    with pytest.raises(RuntimeError):
        ops.linear(torch.zeros(2, 4), torch.zeros(3, 4))
    if not torch.cuda.is_available():
        from renderih_b200.manolayer import ManoLayer
        layer = ManoLayer(A.synthetic_mano(0))
        with pytest.raises(RuntimeError):
            layer(torch.eye(3)[None], torch.zeros(1, 45), torch.zeros(1, 10))


def test_config_and_hrnet_gate(tmp_path):
    from renderih_b200.config import load_cfg
    from renderih_b200.model import load_model
    cfg = load_cfg()
    assert cfg.MODEL.GCN_IN_DIM == [512, 256, 128] and cfg.TRAIN.dropout == 0.05 and cfg.SEED == 88
    p = tmp_path / 'c.yaml'
    p.write_text('MODEL:\n  ENCODER_TYPE: hrnet48\nTRAIN:\n  dropout: 0.0\n')
    cfg2 = load_cfg(str(p))
    assert cfg2.TRAIN.dropout == 0.0 and cfg2.MODEL.graph_k == 2
    m = load_model(cfg2, assets=A.synthetic_assets(0))        # HRNet-w48 encoder (BASELINE config 5)
    sd = m.state_dict()
    assert len(sd) == 2693 and sum(p.numel() for p in m.parameters()) == 88023274      # reference: 2693 keys, 88.02 M parameters
    assert sd['encoder.hrnet.stage4.2.fuse_layers.3.0.2.0.weight'].shape == (384, 48, 3, 3)
    assert sd['mid_model.downsamp_modules.2.0.bias'].shape == (1024,) and sd['encoder.hms_decoder.0.weight'].shape == (720, 720, 1, 1)
    assert m.mid_model.get_info() == {'global_feature_dim': 2048, 'fmaps_dim': [256, 256, 256, 256]}
    cfg2.MODEL.ENCODER_TYPE = 'vit'
    with pytest.raises(NotImplementedError):
        load_model(cfg2, assets=A.synthetic_assets(0))


def test_graph_csr_matches_dense():
    from renderih_b200.model import GraphCSR
    L = A.synthetic_assets(1)['left_graph']['coarsen_graphs_L'][3]
    g = GraphCSR(L)
    d = g.dense().numpy()
    assert np.allclose(d, np.asarray(L.todense(), dtype=np.float32))
    rp, ci, va, rpt, cit, vat = g._host
    dt = np.zeros_like(d)
    for r in range(g.V):
        dt[r, cit[rpt[r]:rpt[r + 1]]] = vat[rpt[r]:rpt[r + 1]]
    assert np.allclose(dt, d.T)
    assert (np.diff(rp) <= 16).all()


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from renderih_b200.train import FlatParams
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(8, 4, 3)
    conv.weight.data = conv.weight.data.contiguous(memory_format=torch.channels_last)
    lin = torch.nn.Linear(5, 3)
    params = list(conv.parameters()) + list(lin.parameters())
    before = [p.detach().clone() for p in params]
    fp = FlatParams(params)
    for p, b in zip(params, before):
        assert torch.equal(p, b) and p.stride() == b.stride()           # re-homing preserves values and strides
    fp.zero_grad()
    x = torch.full((2, 8, 6, 6), float(rank + 1))
    (conv(x).sum() + lin(torch.full((2, 5), float(rank + 1))).sum()).backward()
    local = fp.grad.clone()
    assert params[0].grad.data_ptr() == fp.grad.data_ptr()               # autograd accumulated in place into the flat buffer
    w = fp.all_reduce()
    assert w == world
    out[rank] = (local, fp.grad.clone())
    dist.destroy_process_group()


def test_flat_gradient_allreduce_two_ranks_gloo():
    """N>1 path: the ONE exchange of a step is a sum all-reduce of the flat gradient buffer."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    l0, r0 = out[0]
    l1, r1 = out[1]
    assert torch.allclose(r0, l0 + l1) and torch.allclose(r1, l0 + l1)
    assert not torch.allclose(l0, l1)


def test_lr_schedule_matches_reference_scheduler():
    """`lr_at_epoch` against the learning rates the reference's own StepLR_withWarmUp produced (tests/golden/lr_schedule.json, generated
    by stepping the unmodified utils/lr_sc.py scheduler 300 times with the trainer's arguments, core/lijun_trainer.py:147-153)."""
    import json
    from renderih_b200.train import lr_at_epoch
    g = json.load(open(os.path.join(GOLD, 'lr_schedule.json')))
    for i, ref in enumerate(g['lrs']):
        mine = lr_at_epoch(i + 1, g['base_lr'], init_lr=g['init_lr'], warm_up_epoch=g['warm_up_epoch'], gamma=g['gamma'],
                           step_size=g['step_size'], min_thres=g['min_thres'])
        assert abs(mine - ref) <= 1e-12 + 1e-9 * ref, (i + 1, mine, ref)


def test_mano_loss_matches_reference_golden():
    """renderih_b200.loss.ManoLoss / mano_loss_GCN (pure torch, runs on any device) against core/Loss_mano.mano_loss_GCN of the unmodified
    reference (tests/golden/mano_loss_synth.pt): total, every term and the gradient w.r.t. every prediction, below and above NORM_EPOCH."""
    from oracle import fixtures
    from renderih_b200.config import load_cfg
    from renderih_b200.loss import ManoLoss, mano_loss_GCN
    gold = torch.load(os.path.join(GOLD, 'mano_loss_synth.pt'), weights_only=False)
    cfg = load_cfg(None)
    ml, mr = A.synthetic_mano(0, 'left'), A.synthetic_mano(0, 'right')
    jl = torch.from_numpy(np.asarray(ml['J_regressor'].todense(), dtype=np.float32))
    jr = torch.from_numpy(np.asarray(mr['J_regressor'].todense(), dtype=np.float32))
    gl, gr = ManoLoss(jl, ml['f'], 4, 'cpu'), ManoLoss(jr, mr['f'], 4, 'cpu')
    for epoch, g in gold.items():
        pred, lab = fixtures.make_mano_loss_case(2)
        pred = {k: v.clone().requires_grad_(True) for k, v in pred.items()}
        result, params, hlist, other = fixtures.mano_loss_inputs(pred)
        z = torch.zeros(2, 21, 3)
        total, _, terms, _ = mano_loss_GCN(cfg, epoch, gl, gr, None, None, result, params, hlist, other, None, None, None,
                                           lab['v2d_l'], z[..., :2], lab['v2d_r'], z[..., :2], lab['v3d_l'], z, lab['v3d_r'], z,
                                           lab['root_rel'], 256, lab['lp_gt'], lab['ls_gt'], lab['rp_gt'], lab['rs_gt'])
        assert abs(float(total) - g['total']) < 1e-5 * abs(g['total']), (epoch, float(total), g['total'])
        for k, v in g['terms'].items():
            assert abs(float(terms[k]) - v) <= 1e-5 * abs(v) + 1e-9, (epoch, k, float(terms[k]), v)
        total.backward()
        for k, v in g['grads'].items():
            d = (pred[k].grad - v).abs().max() / v.abs().max().clamp_min(1e-12)
            assert float(d) < 1e-5, (epoch, k, float(d))


def test_input_host_helpers_match_reference_golden():
    """Host-side mirrors of the loader code (renderih_b200/input.py): get_affine_mat == imgUtils.get_affine_mat bit for bit (through the
    oracle, which is pinned to the reference's process_data golden), the cv::warpAffine matrix inversion, and prepare_labels against the
    labels the reference's handDataset.process_data produced (tests/golden/augment_synth.pt)."""
    import numpy as np
    import torch
    from oracle import augment_ref, fixtures
    from renderih_b200 import input as I
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'augment_synth.pt'), weights_only=False)
    S = gold['samples']
    rng = np.random.RandomState(0)
    for _ in range(20):
        th, sc, u, v = rng.uniform(-90, 90), rng.uniform(0.5, 1.5), rng.uniform(-20, 20), rng.uniform(-20, 20)
        M = I.get_affine_mat(th, sc, u, v, 256, 256)
        assert M.dtype == np.float32 and np.array_equal(M, augment_ref.get_affine_mat(th, sc, u, v, 256, 256))
        assert np.array_equal(I._invert_affine(M[0:2]), augment_ref.invert_affine(M[0:2]))
    _, dicts = fixtures.make_augment_case(3)
    hd = {s: {k: torch.from_numpy(np.stack([d[s][k] for d in dicts])) for k in dicts[0][s]} for s in ('left', 'right')}
    A = np.stack([I.get_affine_mat(s['theta'], s['scale'], s['u'], s['v'], 256, 256) for s in S])
    lab = I.prepare_labels(hd, [s['theta'] for s in S], A, [s['flip'] for s in S], bone_length=gold['meta']['bone_length'])
    assert list(lab.keys()) == ['root_rel', 'v2d_l', 'v2d_r', 'v3d_l', 'v3d_r', 'j2d_l', 'j2d_r', 'j3d_l', 'j3d_r']
    for k in lab:
        ref = torch.stack([s['labels'][k] for s in S])
        assert float((lab[k] - ref).abs().max()) <= 2e-6 * float(ref.abs().max()), k


def test_metrics_jr_matches_oracle():
    import numpy as np
    import torch
    from oracle import metrics_ref
    from renderih_b200 import assets as A
    from renderih_b200.metrics import Jr, SAMPLE_FIELDS
    J16 = torch.from_numpy(np.asarray(A.synthetic_mano(0, 'left')['J_regressor'].todense(), dtype='float32'))
    jr = Jr(J16, device='cpu')
    assert torch.equal(jr.J_regressor, metrics_ref.joint_regressor21(J16))
    v = torch.randn(3, 778, 3)
    assert torch.allclose(jr(v), torch.matmul(metrics_ref.joint_regressor21(J16), v))
    assert len(SAMPLE_FIELDS) == 8


def test_mano_pose_helpers_match_reference_golden():
    """rotations.py / ManoLayer conversion helpers vs values the unmodified reference produced (oracle/make_golden.py mano_helpers):
    models/manolayer.py:20-98 (vec2mat, rodrigues_batch, build_mano_frame), 163-248 (pca/axis/Rmat conversions, get_local_frame, SE3)."""
    from renderih_b200 import rotations as R
    from renderih_b200.manolayer import ManoLayer
    gold = torch.load(os.path.join(GOLD, 'mano_helpers_synth.pt'), weights_only=False)
    i, o = gold['inputs'], gold['outputs']
    layer = ManoLayer(A.synthetic_mano(0, 'right'), center_idx=9, use_pca=True)
    Rm = R.rodrigues_batch(i['rot_axis'])
    se3 = layer.buildSE3_batch(R.rodrigues_batch(i['rot_axis'][:5]), i['t'])
    got = {'rodrigues': R.rodrigues_batch(i['axis']), 'vec2mat': R.vec2mat(i['vec6']), 'frame': R.build_mano_frame(i['skel']),
           'Rmat2axis': layer.Rmat2axis(Rm), 'pca2axis': layer.pca2axis(i['pca']), 'pca2Rmat': layer.pca2Rmat(i['pca']),
           'axis2pca': layer.axis2pca(i['axis45']), 'Rmat2pca': layer.Rmat2pca(layer.axis2Rmat(i['axis45'])),
           'local_frame': layer.get_local_frame(i['shape']), 'se3': se3, 'se3_apply': layer.SE3_apply(se3, i['v'])}
    for k, ref in o.items():
        assert got[k].shape == ref.shape, k
        err = float((got[k] - ref).abs().max())
        assert err < 2e-6, (k, err)       # fp32 round-off of two equivalent closed forms (measured <= 2.4e-7)
    assert torch.allclose(R.rodrigues_batch(torch.zeros(2, 3)), torch.eye(3).expand(2, 3, 3))     # zero pose -> identity
