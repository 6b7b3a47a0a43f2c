"""Drive the UNMODIFIED reference (models.model.load_model + core.Loss.calc_loss_GCN + torch.optim.AdamW) exactly as its trainer does
(core/lijun_trainer.py:131-144, 262-313: forward, loss, zero_grad / backward / step) on CPU or CUDA tensors.

TEST / BASELINE INFRASTRUCTURE -- used by tests/ (parity of the product's TrainStep against one reference step), by
`bench.py --impl reference` / `cpu_baseline` (the reference's own CPU path timed beside ours) and by bench.py's
`gpu_eager_baseline` (the reference graph `.cuda()`, eager, the comparator BASELINE.json's north_star names).  Never imported
by the product package.  The reference code that runs is the reference's own (oracle/ref_bridge.py imports it in place).
"""
import os
import tempfile

import torch

from . import fixtures, ref_bridge as rb


def available():
    return rb.reference_available()


class ReferenceStep:
    """model = models.model.load_model(cfg) (ResNet-50 / HRNet cfg) with seeded weights (oracle/fixtures.init_state_dict, the same
    initialisation every product test uses), synthetic or real graph assets; loss = core.Loss.calc_loss_GCN with GraphLoss built like
    core/lijun_trainer.py:198-215 builds it; optimizer = torch.optim.AdamW(lr=cfg.TRAIN.LR, weight_decay=cfg.TRAIN.weight_decay)."""

    def __init__(self, device='cpu', encoder_type='resnet50', dropout=0.05, real_assets=False, lr=None, weight_decay=None):
        from oracle.make_golden import write_synthetic_asset_dir
        self.ns = rb.import_reference()
        self._tmp = tempfile.TemporaryDirectory()
        if real_assets:
            asset_dir = rb.ASSET_DIR
            assert os.path.exists(os.path.join(asset_dir, 'graph_left.pkl')), 'real assets not staged (python -m oracle.build_ref)'
        else:
            asset_dir = self._tmp.name
            write_synthetic_asset_dir(asset_dir, 0)
        self.asset_dir = asset_dir
        self.model, self.cfg = rb.build_reference_model(asset_dir=asset_dir, encoder_type=encoder_type, dropout=dropout)
        self.model.load_state_dict(fixtures.init_state_dict(self.model.state_dict()))
        self.device = torch.device(device)
        self.model.to(self.device)
        for p in self.model.parameters():
            p.requires_grad_(True)
        self.model.decoder.unsample_layer.weight.requires_grad_(False)      # MODEL.freeze_upsample (core/lijun_trainer.py:115-116)
        ns = self.ns
        mano = {s: ns.mano.ManoLayer(os.path.join(asset_dir, 'mano', 'MANO_%s.pkl' % s.upper()), center_idx=None) for s in ('left', 'right')}
        self.gl = ns.loss.GraphLoss(mano['left'].J_regressor, mano['left'].get_faces(), level=4, device=self.device)
        self.gr = ns.loss.GraphLoss(mano['right'].J_regressor, mano['right'].get_faces(), level=4, device=self.device)
        self.lr = self.cfg.TRAIN.LR if lr is None else lr
        self.wd = self.cfg.TRAIN.weight_decay if weight_decay is None else weight_decay
        self.opt = torch.optim.AdamW([p for p in self.model.parameters() if p.requires_grad], lr=self.lr, weight_decay=self.wd)

    def state_dict(self):
        return {k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()}

    def loss(self, out, labels):
        B = labels['v3d_l'].shape[0]
        z = torch.zeros(B, 21, 3, device=self.device)
        conv = self.model.decoder.converter
        return self.ns.loss.calc_loss_GCN(self.cfg, 0, self.gl, self.gr, conv['left'], conv['right'], out[0], out[1], out[2], out[3],
                                          None, None, None, labels['v2d_l'], z[..., :2], labels['v2d_r'], z[..., :2],
                                          labels['v3d_l'], z, labels['v3d_r'], z, labels['root_rel'], 256, upsample_weight=None)[0]

    def forward_backward(self, img, labels):
        self.opt.zero_grad(set_to_none=True)
        out = self.model(img)
        loss = self.loss(out, labels)
        loss.backward()
        return loss.detach(), out

    def step(self, img, labels):
        """core/lijun_trainer.py:262-313 for one batch."""
        loss, _ = self.forward_backward(img, labels)
        self.opt.step()
        return loss


def time_reference(device, batch, steps, warmup, encoder_type='resnet50', forward_only=False, threads=None, cudnn_tf32=True, budget_s=1e9):
    """Wall / device time per step of the unmodified reference on `device`.  -> dict(ms_per_step, steps, ...).
    CUDA: eager PyTorch, cudnn.benchmark on, CUDA events (SURVEY 8d 'Reference GPU timing').  CPU: perf_counter, `threads` torch threads."""
    import time
    dev = torch.device(device)
    if threads:
        torch.set_num_threads(int(threads))
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    if dev.type == 'cuda':
        torch.backends.cudnn.allow_tf32 = bool(cudnn_tf32)
        torch.backends.cudnn.benchmark = True
    try:
        ref = ReferenceStep(dev, encoder_type=encoder_type)
        g = torch.Generator().manual_seed(fixtures.SEED)
        img = torch.randn(batch, 3, 256, 256, generator=g).to(dev)
        labels = {k: v.to(dev) for k, v in fixtures.make_labels(batch).items()}
        if forward_only:
            ref.model.eval()

            def one():
                with torch.no_grad():
                    return ref.model(img)
        else:
            ref.model.train()

            def one():
                return ref.step(img, labels)
        t_begin = time.perf_counter()
        for _ in range(warmup):
            one()
        done = 0
        if dev.type == 'cuda':
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                one(); done += 1
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / done
        else:
            t0 = time.perf_counter()
            for _ in range(steps):
                one(); done += 1
                if time.perf_counter() - t_begin > budget_s:
                    break
            ms = (time.perf_counter() - t0) / done * 1e3
        return {'ms_per_step': ms, 'steps': done, 'warmup': warmup, 'batch': batch, 'images_per_s': batch / ms * 1e3}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = saved
