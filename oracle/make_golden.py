"""Generate tests/golden/*.pt by running the UNMODIFIED reference (imported in place from /root/reference).

Run in the authoring container only:   python -m oracle.make_golden
The reference's own tests contain no golden vectors for this path; these fixtures, produced
by the reference code itself on seeded synthetic inputs, are what pins the oracle (tests/test_oracle_golden.py)
and, through it, the CUDA path.  Synthetic assets are used so that nothing derived from the licence-restricted
MANO files is committed.
"""
import os
import pickle
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fixtures, ref_bridge as rb   # noqa: E402
from renderih_b200 import assets as rih_assets  # noqa: E402  (asset *generator* only: no kernels involved)

GOLD = os.path.join(ROOT, 'tests', 'golden')


def write_synthetic_asset_dir(tmp, seed=0):
    a = rih_assets.synthetic_assets(seed)
    for side in ('left', 'right'):
        with open(os.path.join(tmp, 'graph_%s.pkl' % side), 'wb') as f:
            pickle.dump(a[side + '_graph'], f)
    with open(os.path.join(tmp, 'v_color.pkl'), 'wb') as f:
        pickle.dump(a['dense_coor'], f)
    with open(os.path.join(tmp, 'upsample.pkl'), 'wb') as f:
        pickle.dump(a['upsample'], f)
    os.makedirs(os.path.join(tmp, 'mano'), exist_ok=True)
    for side in ('left', 'right'):
        with open(os.path.join(tmp, 'mano', 'MANO_%s.pkl' % side.upper()), 'wb') as f:
            pickle.dump(rih_assets.synthetic_mano(seed, side), f)
    return a


def flat(out):
    result, params, hlist, other = out
    d = {}
    for side in ('left', 'right'):
        d['verts3d_' + side] = result['verts3d'][side]
        d['verts2d_' + side] = result['verts2d'][side]
        d['scale_' + side] = params['scale'][side]
        d['trans2d_' + side] = params['trans2d'][side]
        d['v3c_' + side] = hlist[0]['verts3d'][side]
        d['v2c_' + side] = hlist[0]['verts2d'][side]
        if other['verts3d_MANO_list'][side]:
            d['v3list_' + side] = other['verts3d_MANO_list'][side][0]
            d['v2list_' + side] = other['verts2d_MANO_list'][side][0]
    for k in (('hms', 'mask', 'dense') if 'hms' in other else ()):     # the myhand graph variant has no auxiliary maps
        t = other[k] if other[k].dim() == 4 else other[k][:, None]     # HRnet_encoder returns mask as [B,64,64] (encoder.py:235)
        d[k + '_sub'] = t[:, :, ::8, ::8].contiguous()
        d[k + '_mean'] = t.mean(dim=(2, 3))
    return {k: v.detach().clone() for k, v in d.items()}


def model_golden(ns, tmp, encoder_type, fname, bn_probe, B=2):
    """Eval forward + train-mode forward/backward (dropout 0) of the unmodified reference model with `encoder_type`
    ('graph' = the common/myhand default variant, built through oracle/ref_bridge.build_reference_myhand_model)."""
    if encoder_type == 'graph':
        ref, _ = rb.build_reference_myhand_model(tmp, 'graph', dropout=0.0)
        _, cfg = rb.build_reference_model(asset_dir=tmp, encoder_type='resnet50', dropout=0.0)    # cfg for calc_loss_GCN only
    else:
        ref, cfg = rb.build_reference_model(asset_dir=tmp, encoder_type=encoder_type, dropout=0.0)
    sd = fixtures.init_state_dict(ref.state_dict())
    ref.load_state_dict(sd)
    img = fixtures.make_image(B)
    ref.eval()
    with torch.no_grad():
        out_eval = flat(ref(img))
    gold = {'weights_sha256': fixtures.checksum(sd), 'batch': B, 'seed': fixtures.SEED, 'torch': torch.__version__,
            'eval': out_eval}
    # training-mode forward + calc_loss_GCN backward, through the reference's own loss code
    ref.train()
    for p in ref.parameters():
        p.requires_grad_(True)
    ref.decoder.unsample_layer.weight.requires_grad_(False)   # freeze_upsample (core/lijun_trainer.py:115-116)
    out = ref(img)
    labels = fixtures.make_labels(B)
    manoL = ns.mano.ManoLayer(os.path.join(tmp, 'mano', 'MANO_LEFT.pkl'), center_idx=None)
    manoR = ns.mano.ManoLayer(os.path.join(tmp, 'mano', 'MANO_RIGHT.pkl'), center_idx=None)
    gl = ns.loss.GraphLoss(manoL.J_regressor, manoL.get_faces(), level=4, device='cpu')
    gr = ns.loss.GraphLoss(manoR.J_regressor, manoR.get_faces(), level=4, device='cpu')
    z = torch.zeros(B, 21, 3)
    loss, _, mano_d, coarse_d = ns.loss.calc_loss_GCN(
        cfg, 0, gl, gr, ref.decoder.converter['left'], ref.decoder.converter['right'],
        out[0], out[1], out[2], out[3], None, None, None,
        labels['v2d_l'], z[..., :2], labels['v2d_r'], z[..., :2], labels['v3d_l'], z, labels['v3d_r'], z,
        labels['root_rel'], 256, upsample_weight=None)
    loss.backward()
    grads = {}
    for k, p in ref.named_parameters():
        if p.grad is None:
            continue
        g = p.grad
        grads[k] = {'norm': float(g.norm()), 'sum': float(g.sum())}
        if g.numel() <= 4096:
            grads[k]['full'] = g.detach().clone()
    bn1 = getattr(ref.encoder, bn_probe).bn1
    gold['train'] = {'out': flat(out), 'loss': float(loss), 'grads': grads,
                     'bn1_running_mean': bn1.running_mean.detach().clone(),
                     'bn1_running_var': bn1.running_var.detach().clone(),
                     'no_grad_keys': [k for k, p in ref.named_parameters() if p.grad is None]}
    gold['encoder_type'] = encoder_type
    torch.save(gold, os.path.join(GOLD, fname))
    print('%s golden: loss %.6f, %d grads, eval |v3d_l| max %.4f' % (encoder_type, float(loss), len(grads), float(out_eval['verts3d_left'].abs().max())))


def mano_golden(ns, tmp):
    # ---------------- ManoLayer goldens on the synthetic MANO tensors
    mg = {'cases': []}
    inp = fixtures.make_mano_inputs(5)
    root = ns.mano.rodrigues_batch(inp['axis'])
    for side in ('left', 'right'):
        path = os.path.join(tmp, 'mano', 'MANO_%s.pkl' % side.upper())
        for cfgc in ({'center_idx': 9, 'use_pca': True, 'new_skel': False, 'ncomps': 45, 'ts': True},
                     {'center_idx': None, 'use_pca': True, 'new_skel': True, 'ncomps': 30, 'ts': False},
                     {'center_idx': 0, 'use_pca': False, 'new_skel': False, 'ncomps': 0, 'ts': True}):
            layer = ns.mano.ManoLayer(path, center_idx=cfgc['center_idx'], use_pca=cfgc['use_pca'], new_skel=cfgc['new_skel'])
            if cfgc['use_pca']:
                pose = inp['pose_pca'][:, :cfgc['ncomps']]
            else:
                pose = layer.axis2Rmat(inp['pose_axis'])
            tr, sc = (inp['trans'], inp['scale']) if cfgc['ts'] else (None, None)
            v, j = layer(root, pose, inp['shape'], tr, sc)
            mg['cases'].append({'side': side, 'cfg': cfgc, 'v': v.clone(), 'j': j.clone()})
    mg['rodrigues'] = root.clone()
    torch.save(mg, os.path.join(GOLD, 'mano_synth.pt'))
    print('mano golden: %d cases' % len(mg['cases']))


MANO_GRAD_CASES = ({'center_idx': 9, 'use_pca': True, 'new_skel': False, 'ncomps': 45, 'ts': True},
                   {'center_idx': None, 'use_pca': True, 'new_skel': True, 'ncomps': 30, 'ts': True},
                   {'center_idx': 0, 'use_pca': False, 'new_skel': True, 'ncomps': 0, 'ts': False})


def mano_grad_golden(ns, tmp):
    """Gradients of the UNMODIFIED reference ManoLayer (torch autograd) for loss = <v, wv> + <j, wj> on the synthetic MANO tensors."""
    inp = fixtures.make_mano_inputs(5)
    wv, wj = fixtures.make_mano_loss_weights(5)
    out = {'cases': []}
    for side in ('left', 'right'):
        path = os.path.join(tmp, 'mano', 'MANO_%s.pkl' % side.upper())
        for cfgc in MANO_GRAD_CASES:
            layer = ns.mano.ManoLayer(path, center_idx=cfgc['center_idx'], use_pca=cfgc['use_pca'], new_skel=cfgc['new_skel'])
            root = ns.mano.rodrigues_batch(inp['axis']).clone().requires_grad_(True)
            pose = (inp['pose_pca'][:, :cfgc['ncomps']] if cfgc['use_pca'] else layer.axis2Rmat(inp['pose_axis'])).clone().requires_grad_(True)
            shape = inp['shape'].clone().requires_grad_(True)
            tr = inp['trans'].clone().requires_grad_(True) if cfgc['ts'] else None
            sc = inp['scale'].clone().requires_grad_(True) if cfgc['ts'] else None
            v, j = layer(root, pose, shape, tr, sc)
            ((v * wv).sum() + (j * wj).sum()).backward()
            out['cases'].append({'side': side, 'cfg': cfgc, 'd_root': root.grad.clone(), 'd_pose': pose.grad.clone(), 'd_shape': shape.grad.clone(),
                                 'd_trans': None if tr is None else tr.grad.clone(), 'd_scale': None if sc is None else sc.grad.clone()})
    torch.save(out, os.path.join(GOLD, 'mano_grad_synth.pt'))
    print('mano gradient golden: %d cases' % len(out['cases']))


def newgraph_golden(ns, tmp):
    """'newgraph' variant (common/myhand/lijun_model_newgraph.load_new_model: graph decoder + ParamRegressor + MANO tail), eval forward and
    train forward/backward under a fixed linear functional of the outputs (the reference's mano_loss_GCN is not importable here)."""
    ref, _ = rb.build_reference_myhand_model(tmp, 'newgraph', dropout=0.0)
    sd = fixtures.init_state_dict(ref.state_dict())
    ref.load_state_dict(sd)
    img = fixtures.make_image(B)
    with ref._rih_cpu_shims():
        ref.eval()
        with torch.no_grad():
            out_eval = {k: v.detach().clone() for k, v in fixtures.flat_newgraph(ref(img)).items()}
        ref.train()
        for p in ref.parameters():
            p.requires_grad_(True)
        ref.decoder.unsample_layer.weight.requires_grad_(False)
        fo = fixtures.flat_newgraph(ref(img))
        loss = fixtures.newgraph_loss(fo, fixtures.make_newgraph_cotangents(B))
        loss.backward()
    grads = {}
    for k, p in ref.named_parameters():
        if p.grad is None:
            continue
        grads[k] = {'norm': float(p.grad.norm()), 'sum': float(p.grad.sum())}
        if p.grad.numel() <= 4096:
            grads[k]['full'] = p.grad.detach().clone()
    gold = {'weights_sha256': fixtures.checksum(sd), 'batch': B, 'seed': fixtures.SEED, 'torch': torch.__version__, 'eval': out_eval,
            'train': {'out': {k: v.detach().clone() for k, v in fo.items()}, 'loss': float(loss), 'grads': grads,
                      'no_grad_keys': [k for k, p in ref.named_parameters() if p.grad is None]}}
    torch.save(gold, os.path.join(GOLD, 'model_newgraph_synth_b2.pt'))
    print('newgraph golden: loss %.6f, %d grads, %d keys' % (float(loss), len(grads), len(sd)))


def mano_loss_golden(ns, tmp):
    """core/Loss_mano.mano_loss_GCN (the unmodified reference code) on seeded predictions / labels: total, per-term values and the
    gradient with respect to every prediction tensor."""
    import core.Loss_mano as lm
    lm.get_upsample_path = lambda: os.path.join(tmp, 'upsample.pkl')
    _, cfg = rb.build_reference_model(asset_dir=tmp, encoder_type='resnet50', dropout=0.0)
    manoL = ns.mano.ManoLayer(os.path.join(tmp, 'mano', 'MANO_LEFT.pkl'), center_idx=None)
    manoR = ns.mano.ManoLayer(os.path.join(tmp, 'mano', 'MANO_RIGHT.pkl'), center_idx=None)
    gl = lm.ManoLoss(manoL.J_regressor, manoL.get_faces(), level=4, device='cpu')
    gr = lm.ManoLoss(manoR.J_regressor, manoR.get_faces(), level=4, device='cpu')
    out = {}
    for epoch in (0, 60):          # below / above NORM_EPOCH (edge term off / on)
        pred, lab = fixtures.make_mano_loss_case(2)
        pred = {k: v.clone().requires_grad_(True) for k, v in pred.items()}
        result, params, hlist, other = fixtures.mano_loss_inputs(pred)
        z = torch.zeros(2, 21, 3)
        total, _, terms, _ = lm.mano_loss_GCN(cfg, epoch, gl, gr, None, None, result, params, hlist, other, None, None, None,
                                              lab['v2d_l'], z[..., :2], lab['v2d_r'], z[..., :2], lab['v3d_l'], z, lab['v3d_r'], z,
                                              lab['root_rel'], 256, lab['lp_gt'], lab['ls_gt'], lab['rp_gt'], lab['rs_gt'], upsample_weight=None)
        total.backward()
        out[epoch] = {'total': float(total), 'terms': {k: float(v) for k, v in terms.items()}, 'grads': {k: v.grad.clone() for k, v in pred.items()}}
    torch.save(out, os.path.join(GOLD, 'mano_loss_synth.pt'))
    print('mano_loss golden: totals', {e: round(o['total'], 4) for e, o in out.items()})


def eval_metrics_golden(ns, tmp):
    """Evaluation metrics of apps/eval_interhand.py, produced by EXECUTING the reference's own source: the definitions of
    batch_compute_similarity_transform_torch / compute_similarity_transform(_batch) / get_alignMesh / Jr, the accumulator set-up
    (:268-298), the body of the evaluation loop (:300-438, minus the data loading / network call / timing statements) on two seeded
    batches, and the reductions after it (:441-552) -- all compiled from the file under /root/reference at generation time.
    `pytorch3d.ops.knn_points` (used by utils/eval_metrics.compute_cdev) is absent from this container; a brute-force K=1 stand-in with
    the same return convention (squared distances, indices) is injected, and that is stated in the golden's meta."""
    import ast
    import contextlib
    import io
    import types
    path = os.path.join(rb.REF_ROOT, 'apps', 'eval_interhand.py')
    tree = ast.parse(open(path).read())
    wanted = {'batch_compute_similarity_transform_torch', 'compute_similarity_transform', 'compute_similarity_transform_batch', 'get_alignMesh', 'Jr'}
    defs = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in wanted]
    assert len(defs) == len(wanted)
    env = {'torch': torch, 'np': np}
    exec(compile(ast.Module(body=defs, type_ignores=[]), path, 'exec'), env)

    def knn_points(p1, p2, lengths1=None, lengths2=None, K=1, return_nn=False):
        d2 = ((p1[:, :, None, :] - p2[:, None, :, :]) ** 2).sum(-1)
        dist, idx = d2.min(dim=2)
        return dist[:, :, None], idx[:, :, None], None
    p3d, p3d_ops = types.ModuleType('pytorch3d'), types.ModuleType('pytorch3d.ops')
    p3d_ops.knn_points = knn_points
    p3d.ops = p3d_ops
    sys.modules['pytorch3d'], sys.modules['pytorch3d.ops'] = p3d, p3d_ops
    sys.modules.pop('utils.eval_metrics', None)

    main_if = [n for n in tree.body if isinstance(n, ast.If)][0]
    with_node = [n for n in main_if.body if isinstance(n, ast.With)][0]
    loop = with_node.body[0]
    assert isinstance(loop, ast.For)
    setup = [n for n in main_if.body if 268 <= n.lineno <= 298]
    tail = [n for n in main_if.body if n.lineno > with_node.lineno]

    def keep(stmt):                      # drop data loading (.cuda() of the loader tuple), the network call and the wall-clock statements
        text = ast.unparse(stmt)
        return not (('data[' in text and '.cuda()' in text) or 'network(' in text or 'time.time()' in text or text.startswith('total_time'))
    body = [n for n in loop.body if keep(n)]
    dropped = [ast.unparse(n) for n in loop.body if not keep(n)]
    assert len(dropped) == 9, dropped

    def run(stmts):
        for st in stmts:
            exec(compile(ast.Module(body=[st], type_ignores=[]), path, 'exec'), env)
            if isinstance(st, ast.Assign) and ast.unparse(st.targets[0]) == 'error_mpjpe':
                env.setdefault('_error_mpjpe', []).append(float(env['error_mpjpe']))

    manoL = ns.mano.ManoLayer(os.path.join(tmp, 'mano', 'MANO_LEFT.pkl'), center_idx=None)
    manoR = ns.mano.ManoLayer(os.path.join(tmp, 'mano', 'MANO_RIGHT.pkl'), center_idx=None)
    env['J_regressor'] = {'left': env['Jr'](manoL.J_regressor, device='cpu'), 'right': env['Jr'](manoR.J_regressor, device='cpu')}
    case = fixtures.make_eval_case(6)
    n = case['gt_left'].shape[0]
    env.update({'iou033': np.arange(n) % 3 == 0, 'iou067': np.arange(n) % 3 == 1, 'iou1': np.arange(n) % 3 == 2})
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), torch.no_grad():
        run(setup)
        for lo, hi in ((0, 4), (4, 6)):          # two batches of different size: the reference concatenates them after the loop
            env.update({'verts_left_gt': case['gt_left'][lo:hi].clone(), 'verts_right_gt': case['gt_right'][lo:hi].clone(),
                        'result': {'verts3d': {'left': case['pred_left'][lo:hi].clone(), 'right': case['pred_right'][lo:hi].clone()}}})
            run(body)
        run(tail)
    t = lambda a: torch.as_tensor(np.asarray(a))
    out = {'meta': {'source': 'apps/eval_interhand.py executed in place (defs + loop body :300-438 + reductions :441-552)',
                    'knn_points': 'brute-force stand-in for pytorch3d.ops.knn_points (absent here)', 'torch': torch.__version__,
                    'batches': [4, 2], 'dropped_loop_statements': dropped},
           'per_element': {side: {k: t(env[k][side]) for k in ('orijoint_loss', 'orivert_loss', 'joints_loss', 'verts_loss', 'pajoints_loss', 'paverts_loss')}
                           for side in ('left', 'right')},
           'mrrpe': t(env['mrrpe']), 'cdev': env['error'].clone(),
           'summary_mm': {'ori_mpjpe_left': float(env['orijoint_left']), 'ori_mpjpe_right': float(env['orijoint_right']),
                          'ori_mpvpe_left': float(env['orivert_left']), 'ori_mpvpe_right': float(env['orivert_right']),
                          'mpjpe_left': float(env['joints_loss']['left'].mean() * 1000), 'mpjpe_right': float(env['joints_loss']['right'].mean() * 1000),
                          'mpvpe_left': float(env['verts_loss']['left'].mean() * 1000), 'mpvpe_right': float(env['verts_loss']['right'].mean() * 1000),
                          'pa_mpjpe_left': float(env['joints_mean_loss_left']), 'pa_mpjpe_right': float(env['joints_mean_loss_right']),
                          'pa_mpvpe_left': float(env['verts_mean_loss_left']), 'pa_mpvpe_right': float(env['verts_mean_loss_right']),
                          'double_pa_mpjpe': float(1000 * np.mean(env['pa_joint_error'])), 'double_pa_mpvpe': float(1000 * np.mean(env['pa_mesh_error'])),
                          'double_mpjpe': 1000 * env['_error_mpjpe'][0], 'double_mpvpe': 1000 * env['_error_mpjpe'][1]},
           'double_per_sample': {'pa_joint': t(env['get_alignMesh'](env['pred_3djoint'], env['gt_3djoint'], reduction=None)[0]),
                                 'pa_mesh': t(env['get_alignMesh'](env['pred_mesh'], env['gt_mesh'], reduction=None)[0])},
           'printed': sink.getvalue()}
    torch.save(out, os.path.join(GOLD, 'eval_metrics_synth.pt'))
    print('eval_metrics golden:', {k: round(v, 4) for k, v in out['summary_mm'].items()})


def augment_golden(ns, tmp):
    """Training-time loader path, produced by EXECUTING the reference's own `handDataset.augm_params` / `handDataset.process_data`
    (core/loader.py:96-220, compiled from the file under /root/reference -- the module itself cannot be imported here: imgaug, dataset
    classes) with the reference's `imgUtils` (utils/manoutils.py) and this container's cv2 / torchvision.  The random draws of each
    sample are recorded (augm_params outputs; the `a`, `b` of imgUtils.add_noise replayed from the saved generator states)."""
    import ast
    import random
    import types
    import cv2 as cv
    import torchvision.transforms as transforms
    from utils.manoutils import imgUtils
    from dataset.dataset_utils import BONE_LENGTH
    path = os.path.join(rb.REF_ROOT, 'core', 'loader.py')
    tree = ast.parse(open(path).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'handDataset'][0]
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ('augm_params', 'process_data')]
    assert len(fns) == 2
    env = {'imgUtils': imgUtils, 'cv': cv, 'np': np, 'torch': torch, 'random': random}
    exec(compile(ast.Module(body=fns, type_ignores=[]), path, 'exec'), env)
    me = types.SimpleNamespace(train=True, seq=None, noise=0.0, flip=True, theta=[-90, 90], scale=[0.75, 1.25], uv=[-10, 10], bone_length=BONE_LENGTH,
                               normalize_img=transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]))
    draws = {}
    orig_noise = imgUtils.add_noise

    def recording_noise(img, noise=0.0, scale=255.0, alpha=0.3, beta=0.05):
        st_np, st_py = np.random.get_state(), random.getstate()
        draws['a'] = np.random.uniform(1 - alpha, 1 + alpha, 3)
        draws['b'] = scale * beta * (2 * random.random() - 1)
        np.random.set_state(st_np); random.setstate(st_py)
        return orig_noise(img, noise=noise, scale=scale, alpha=alpha, beta=beta)

    def recording_params():
        out = env['augm_params'](me)
        draws['params'] = out
        return out
    me.augm_params = recording_params
    imgUtils.add_noise = staticmethod(recording_noise)
    frames, dicts = fixtures.make_augment_case(3)
    samples = []
    try:
        for i, (f, hd) in enumerate(zip(frames, dicts)):
            random.seed(100 + i); np.random.seed(100 + i)
            out = env['process_data'](me, f.copy(), {s: {k: v.copy() for k, v in d.items()} for s, d in hd.items()})
            theta, scale, u, v, flip = draws['params']
            names = ('ori_img', 'imgTensor', 'v2d_l', 'j2d_l', 'v2d_r', 'j2d_r', 'v3d_l', 'j3d_l', 'v3d_r', 'j3d_r', 'root_rel')
            rec = dict(zip(names, out))
            u8 = (rec.pop('ori_img') * 255).round().to(torch.uint8).permute(1, 2, 0).contiguous()        # exact: ori_img = uint8 / 255
            samples.append({'theta': theta, 'scale': scale, 'u': u, 'v': v, 'flip': bool(flip), 'a': draws['a'].copy(), 'b': float(draws['b']),
                            'final_u8_bgr': u8, 'imgTensor': rec.pop('imgTensor') if i == 0 else (rec.pop('imgTensor'), None)[1], 'labels': rec})
    finally:
        imgUtils.add_noise = staticmethod(orig_noise)
    torch.save({'meta': {'source': 'core/loader.py handDataset.augm_params / process_data executed in place; utils/manoutils.imgUtils',
                         'cv2': cv.__version__, 'torch': torch.__version__, 'bone_length': BONE_LENGTH}, 'samples': samples},
               os.path.join(GOLD, 'augment_synth.pt'))
    print('augment golden:', [(round(s['theta'], 2), round(s['scale'], 3), s['flip']) for s in samples])


def mano_helpers_golden(ns, tmp):
    """The pose-representation helpers around the ManoLayer (models/manolayer.py:20-98, 163-248): rodrigues_batch, vec2mat,
    build_mano_frame, the pca/axis/Rmat conversions, get_local_frame, buildSE3_batch / SE3_apply -- outputs of the unmodified reference
    on seeded inputs (synthetic MANO tables).  Batch sizes avoid 3: the reference's `torch.cross` calls carry no `dim`, so a batch of
    exactly 3 frames would be crossed along the batch axis (a reference quirk this package does not reproduce)."""
    g = torch.Generator().manual_seed(fixtures.SEED + 7)
    layer = ns.mano.ManoLayer(os.path.join(tmp, 'mano', 'MANO_RIGHT.pkl'), center_idx=9, use_pca=True)
    ax = torch.randn(40, 3, generator=g) * 1.5
    ax[0] = 0
    inp = {'axis': ax, 'vec6': torch.randn(16, 6, generator=g), 'skel': torch.randn(4, 21, 3, generator=g),
           'rot_axis': torch.randn(30, 3, generator=g) * 2.5, 'pca': torch.randn(4, 30, generator=g), 'axis45': torch.randn(4, 45, generator=g) * 0.5,
           'shape': torch.randn(5, 10, generator=g), 't': torch.randn(5, 3, 1, generator=g), 'v': torch.randn(5, 3, generator=g)}
    Rm = ns.mano.rodrigues_batch(inp['rot_axis'])
    R5 = ns.mano.rodrigues_batch(inp['rot_axis'][:5])
    se3 = layer.buildSE3_batch(R5, inp['t'])
    out = {'rodrigues': ns.mano.rodrigues_batch(ax), 'vec2mat': ns.mano.vec2mat(inp['vec6']), 'frame': ns.mano.build_mano_frame(inp['skel']),
           'Rmat2axis': layer.Rmat2axis(Rm), 'pca2axis': layer.pca2axis(inp['pca']), 'pca2Rmat': layer.pca2Rmat(inp['pca']),
           'axis2pca': layer.axis2pca(inp['axis45']), 'Rmat2pca': layer.Rmat2pca(layer.axis2Rmat(inp['axis45'])),
           'local_frame': layer.get_local_frame(inp['shape']), 'se3': se3, 'se3_apply': layer.SE3_apply(se3, inp['v'])}
    torch.save({'inputs': inp, 'outputs': {k: v.detach().clone() for k, v in out.items()}, 'torch': torch.__version__},
               os.path.join(GOLD, 'mano_helpers_synth.pt'))
    print('mano helpers golden:', {k: tuple(v.shape) for k, v in out.items()})


def main(which):
    """which: any of 'resnet50', 'hrnet48', 'graph', 'newgraph', 'mano_loss', 'eval_metrics', 'augment', 'mano', 'mano_grad', 'mano_helpers' (default: all).  Each golden file is written independently."""
    os.makedirs(GOLD, exist_ok=True)
    ns = rb.import_reference()
    with tempfile.TemporaryDirectory() as tmp:
        write_synthetic_asset_dir(tmp, 0)
        if 'resnet50' in which:
            model_golden(ns, tmp, 'resnet50', 'model_synth_b2.pt', 'resnet')
        if 'resnet50_b16' in which:   # same model at batch 16: BatchNorm statistics over >= 1024 samples, so train-mode parity can be held much tighter
            model_golden(ns, tmp, 'resnet50', 'model_synth_b16.pt', 'resnet', B=16)
        if 'hrnet48' in which:     # BASELINE config 5 (models/encoder.py:176-352, model_zoo/hrnet.py)
            model_golden(ns, tmp, 'hrnet48', 'model_hrnet48_synth_b2.pt', 'hrnet')
        if 'graph' in which:       # common/myhand/lijun_model_graph.load_graph_model
            model_golden(ns, tmp, 'graph', 'model_graph_synth_b2.pt', 'resnet')
        if 'newgraph' in which:    # SURVEY 8(f) row 1, MANO tail
            newgraph_golden(ns, tmp)
        if 'mano_loss' in which:   # core/Loss_mano.py:245-335
            mano_loss_golden(ns, tmp)
        if 'augment' in which:        # core/loader.py augmentation + normalisation
            augment_golden(ns, tmp)
        if 'eval_metrics' in which:   # apps/eval_interhand.py metric loop
            eval_metrics_golden(ns, tmp)
        if 'mano' in which:
            mano_golden(ns, tmp)
        if 'mano_grad' in which:
            mano_grad_golden(ns, tmp)
        if 'mano_helpers' in which:
            mano_helpers_golden(ns, tmp)


if __name__ == '__main__':
    main(sys.argv[1:] or ['resnet50', 'resnet50_b16', 'hrnet48', 'graph', 'newgraph', 'mano_loss', 'eval_metrics', 'augment', 'mano', 'mano_grad', 'mano_helpers'])
