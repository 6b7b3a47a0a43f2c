"""Training step plumbing for the hot path: flat parameter / gradient buffers, fused AdamW, one NCCL all-reduce,
whole-step CUDA-graph capture.

Mirrors what the reference's trainer does around `network(imgTensors)` (core/lijun_trainer.py:262-313: forward,
calc_loss_GCN, zero_grad / backward / AdamW step, DDP gradient all-reduce) with B200-first mechanics:
  * every trainable tensor is a view of ONE flat fp32 buffer, gradients likewise -> zero_grad is one memset,
    the data-parallel exchange is ONE all-reduce of the flat gradient (155 MB) over NVLink, AdamW is one kernel;
  * BatchNorm statistics stay per rank (the reference uses no SyncBN);
  * the whole step (forward, loss, backward, optimizer) is captured into a CUDA graph and replayed.
"""
import torch
import torch.distributed as dist

from ._lib import call
from . import ops


def lr_at_epoch(epoch, base_lr, init_lr=None, warm_up_epoch=3, gamma=0.1, step_size=80, min_thres=0.05):
    """Learning rate of the reference's `StepLR_withWarmUp` (utils/lr_sc.py:159-175) at `epoch` (= the scheduler's `last_epoch`):
    linear warm-up from init_lr (the trainer passes 1e-2 * LR, core/lijun_trainer.py:147-153) over `warm_up_epoch` epochs, then
    base_lr * max(gamma ** ((epoch - warm_up) // step_size), min_thres)."""
    if init_lr is None:
        init_lr = 1e-2 * base_lr
    if epoch < warm_up_epoch:
        return init_lr + (base_lr - init_lr) * (epoch / warm_up_epoch)
    return base_lr * max(gamma ** ((epoch - warm_up_epoch) // step_size), min_thres)


class FlatParams:
    """Re-home the given parameters (and their .grad) into flat contiguous buffers, preserving each tensor's strides
    (conv weights stay channels_last)."""

    def __init__(self, params):
        params = [p for p in params if p.requires_grad]
        assert params, 'no trainable parameters'
        dev = params[0].device
        self.params = params
        total = 0
        offs = []
        for p in params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4        # keep every tensor 16-byte aligned
        self.numel = total
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        self.step_count = 0
        for p, off in zip(params, offs):
            n = p.numel()
            assert p.is_contiguous() or (p.dim() == 4 and p.permute(0, 2, 3, 1).is_contiguous()), 'dense tensors only'
            view = torch.as_strided(self.flat, p.shape, p.stride(), off)
            view.copy_(p.data)
            p.data = view
            p.grad = torch.as_strided(self.grad, p.shape, p.stride(), off)
            if p.is_cuda:
                ops.register_grad_target(p, p.grad)   # backward kernels accumulate straight into the flat gradient

    def zero_grad(self):
        self.grad.zero_()

    def all_reduce(self, group=None):
        """The single gradient exchange of a data-parallel step (sum; the mean is folded into adamw_step)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
            return dist.get_world_size(group)
        return 1

    def adamw_step(self, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_scale=1.0, step=None):
        """torch.optim.AdamW semantics (decoupled weight decay) as one kernel over the flat buffer."""
        if step is None:
            self.step_count += 1
            step = self.step_count
        call('rih_adamw_step', self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
             self.numel, float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(step), float(grad_scale),
             torch.cuda.current_stream().cuda_stream)


def trainable_used_params(model, loss_fn, img):
    """Parameters that actually receive a gradient (the reference needs find_unused_parameters=True because 63 tensors
    never do): found with one dry-run backward."""
    for p in model.parameters():
        p.grad = None
    out = model(img)
    loss_fn(out).backward()
    used = [p for p in model.parameters() if p.requires_grad and p.grad is not None]
    for p in model.parameters():
        p.grad = None
    return used


class TrainStep:
    """One data-parallel training step: forward -> loss -> backward -> (all-reduce) -> AdamW, optionally CUDA-graphed."""

    def __init__(self, model, loss_fn, example_img, lr=3e-4, weight_decay=1e-2, use_graph=True, group=None):
        self.model, self.loss_fn, self.group = model, loss_fn, group
        self.lr, self.wd = lr, weight_decay
        self.base_lr = lr
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.flatp = FlatParams(trainable_used_params(model, loss_fn, example_img))
        try:      # the two decoder streams make some AccumulateGrad nodes run on a side stream; torch's advisory warning is expected here
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except Exception:
            pass
        self.static_img = example_img.clone()
        self.graph = None
        self.loss = None
        self.use_graph = use_graph

    def set_epoch(self, epoch, cfg=None, **kw):
        """Learning-rate schedule of the reference trainer: the fused AdamW kernel takes the rate as a launch argument
        (it runs outside the captured graph), so the schedule costs nothing per step."""
        if cfg is not None:
            kw = dict(warm_up_epoch=cfg.TRAIN.warm_up, gamma=cfg.TRAIN.lr_decay_gamma, step_size=cfg.TRAIN.lr_decay_step, **kw)
            self.base_lr = cfg.TRAIN.LR
        self.lr = lr_at_epoch(epoch, getattr(self, 'base_lr', self.lr), **kw)
        return self.lr

    def _step_body(self):
        self.flatp.zero_grad()
        out = self.model(self.static_img)
        loss = self.loss_fn(out)
        loss.backward()
        return loss.detach()

    def _eager(self):
        loss = self._step_body()
        self.flatp.all_reduce(self.group)
        self.flatp.adamw_step(self.lr, weight_decay=self.wd, grad_scale=1.0 / self.world)
        return loss

    def capture(self, warmup=3):
        """Warm up on a side stream, then capture fwd+bwd (graph 1) and the optimizer (graph 2); the NCCL all-reduce
        runs between the two replays on the same stream (single-GPU: nothing in between)."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._step_body()
        self.flatp.step_count += 0
        return self

    def prefetch(self, img_host):
        """Start the host -> device copy of the NEXT step's batch (pinned host tensor) on a copy stream, overlapping the current step; the
        next `__call__()` (without an argument) moves it into the graph's static input with one device-to-device copy.  This is the
        input double-buffering a DataLoader with `pin_memory` + `non_blocking` gives the reference trainer (core/lijun_trainer.py:255-262)."""
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.static_img.device)
            self._stage = torch.empty_like(self.static_img)
            self._stage_free = None
        if self._stage_free is not None:
            self._copy_stream.wait_event(self._stage_free)      # the previous stage -> static copy must have consumed the buffer
        with torch.cuda.stream(self._copy_stream):
            self._stage.copy_(img_host, non_blocking=True)
        self._staged = True

    def __call__(self, img=None):
        if img is not None:
            self.static_img.copy_(img, non_blocking=True)
        elif getattr(self, '_staged', False):
            cur = torch.cuda.current_stream()
            cur.wait_stream(self._copy_stream)
            self.static_img.copy_(self._stage, non_blocking=True)
            self._stage_free = torch.cuda.Event()
            self._stage_free.record(cur)
            self._staged = False
        if self.graph is None:
            return self._eager()
        self.graph.replay()
        self.flatp.all_reduce(self.group)
        self.flatp.adamw_step(self.lr, weight_decay=self.wd, grad_scale=1.0 / self.world)
        return self.loss
