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


def model_golden(ns, tmp, encoder_type, fname, bn_probe):
    """Eval forward + train-mode forward/backward (dropout 0) of the unmodified reference model with `encoder_type`
    ('graph' = the common/myhand default variant, built through oracle/ref_bridge.build_reference_myhand_model)."""
    if encoder_type == 'graph':
        ref, _ = rb.build_reference_myhand_model(tmp, 'graph', dropout=0.0)
        _, cfg = rb.build_reference_model(asset_dir=tmp, encoder_type='resnet50', dropout=0.0)    # cfg for calc_loss_GCN only
    else:
        ref, cfg = rb.build_reference_model(asset_dir=tmp, encoder_type=encoder_type, dropout=0.0)
    sd = fixtures.init_state_dict(ref.state_dict())
    ref.load_state_dict(sd)
    B = 2
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
    B = 2
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


def main(which):
    """which: any of 'resnet50', 'hrnet48', 'graph', 'newgraph', 'mano_loss', 'mano', 'mano_grad' (default: all).  Each golden file is written independently."""
    os.makedirs(GOLD, exist_ok=True)
    ns = rb.import_reference()
    with tempfile.TemporaryDirectory() as tmp:
        write_synthetic_asset_dir(tmp, 0)
        if 'resnet50' in which:
            model_golden(ns, tmp, 'resnet50', 'model_synth_b2.pt', 'resnet')
        if 'hrnet48' in which:     # BASELINE config 5 (models/encoder.py:176-352, model_zoo/hrnet.py)
            model_golden(ns, tmp, 'hrnet48', 'model_hrnet48_synth_b2.pt', 'hrnet')
        if 'graph' in which:       # common/myhand/lijun_model_graph.load_graph_model
            model_golden(ns, tmp, 'graph', 'model_graph_synth_b2.pt', 'resnet')
        if 'newgraph' in which:    # SURVEY 8(f) row 1, MANO tail
            newgraph_golden(ns, tmp)
        if 'mano_loss' in which:   # core/Loss_mano.py:245-335
            mano_loss_golden(ns, tmp)
        if 'mano' in which:
            mano_golden(ns, tmp)
        if 'mano_grad' in which:
            mano_grad_golden(ns, tmp)


if __name__ == '__main__':
    main(sys.argv[1:] or ['resnet50', 'hrnet48', 'graph', 'newgraph', 'mano_loss', 'mano', 'mano_grad'])
