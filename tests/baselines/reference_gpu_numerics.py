"""How far is the REFERENCE's own GPU arithmetic from its CPU fp32 result?  Runs the reference-equivalent torch graph
(oracle/model_ref.py) on the GPU with cuDNN TF32 on (the reference's default) and off, eval mode, batch 2, golden weights,
and prints the per-output relative error against the golden produced by the unmodified reference on the CPU."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fixtures, model_ref  # noqa: E402
from renderih_b200 import assets as A   # noqa: E402
from renderih_b200.model import load_model  # noqa: E402

gold = torch.load(os.path.join(ROOT, 'tests', 'golden', 'model_synth_b2.pt'), weights_only=False)
a = A.synthetic_assets(0)
sd = fixtures.init_state_dict(load_model(assets=a).state_dict())
assert fixtures.checksum(sd) == gold['weights_sha256']
sd = {k: v.cuda() for k, v in sd.items()}
Ap = model_ref.prepare_assets(a)
for s in Ap:
    Ap[s]['L'] = [l.cuda() for l in Ap[s]['L']]
img = fixtures.make_image(gold['batch']).cuda()
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        out = model_ref.model_forward({k: v.clone() for k, v in sd.items()}, Ap, img, training=False)
    res = {}
    for side in ('left', 'right'):
        for k, t in (('verts3d_', out[0]['verts3d'][side]), ('verts2d_', out[0]['verts2d'][side]), ('trans2d_', out[1]['trans2d'][side])):
            g = gold['eval'][k + side]
            res[k + side] = float((t.cpu().double() - g.double()).abs().max() / g.double().abs().max())
    g = gold['eval']['hms_mean']
    res['hms_mean'] = float((out[3]['hms'].mean(dim=(2, 3)).cpu().double() - g.double()).abs().max() / g.double().abs().max())
    print('reference torch-GPU eval forward, cudnn.allow_tf32=%s: rel err vs CPU fp32 golden:' % tf32, {k: '%.2e' % v for k, v in res.items()})
