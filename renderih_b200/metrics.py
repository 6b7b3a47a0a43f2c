"""Evaluation metrics of the reference's test loop (apps/eval_interhand.py:147-170, 300-552; utils/eval_metrics.py:36-50) on the GPU:
`Jr`, `batch_metrics` (two kernels per batch, csrc/metrics.cu) and the `EvalMetrics` accumulator that reproduces the numbers the reference
prints (MPJPE / MPVPE raw, bone-length rescaled and Procrustes aligned per hand, "mrrpe", contact deviation, the "double" variants).

Deliberate difference: `batch_compute_similarity_transform_torch` decides its layout from `S1.shape[0]` (:35), which on batched input is
the batch size, so the reference mis-handles batches of exactly 2 or 3 samples; the kernel treats every batch size alike.
There is no CPU path: inputs must be CUDA tensors.
"""
import ctypes

import torch

from ._lib import call

SAMPLE_FIELDS = ('ori_mpjpe', 'ori_mpvpe', 'mpjpe', 'mpvpe', 'pa_mpjpe', 'pa_mpvpe', 'double_pa_mpjpe', 'double_pa_mpvpe')


class Jr:
    """apps/eval_interhand.py:147-170: the 16-joint MANO regressor extended by the 5 finger-tip vertices, reordered to 21 joints."""

    def __init__(self, J_regressor, device='cuda'):
        self.device = device
        J = J_regressor.clone().detach().float()
        tips = torch.zeros_like(J[:5])
        for i, v in enumerate((745, 317, 444, 556, 673)):
            tips[i, v] = 1.0
        J = torch.cat([J, tips], dim=0)
        order = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]
        self.J_regressor = J[order].contiguous().to(device)

    def __call__(self, v):
        return torch.matmul(self.J_regressor, v)


def batch_metrics(J_regressor, pred_left, pred_right, gt_left, gt_right, per_element=True, contact_dist=3e-3):
    """One batch of the evaluation loop.  J_regressor: {'left': Jr, 'right': Jr}; vertices [B,778,3] CUDA fp32 in metres (absolute).
    Returns per-sample tensors named after the reference's lists:
      orijoint_loss / joints_loss {side: [B,21]}, orivert_loss / verts_loss {side: [B,778]}   (per_element=True)
      pajoints_loss / paverts_loss {side: [B]}, mrrpe [B,3], cdev [B] (NaN without contact),
      double_pa_joint / double_pa_mesh / double_joint / double_mesh [B], sample [2,B,8] (means over points, SAMPLE_FIELDS)."""
    tens = [t.contiguous().float() for t in (pred_left, gt_left, pred_right, gt_right)]
    if not all(t.is_cuda for t in tens):
        raise RuntimeError('renderih_b200.metrics.batch_metrics: CUDA tensors required (there is no CPU path)')
    B = tens[0].shape[0]
    for t in tens:
        assert tuple(t.shape) == (B, 778, 3), 'vertices must be [B,778,3]'
    dev = tens[0].device
    Jl, Jr_ = (J_regressor[s].J_regressor.to(dev).float().contiguous() for s in ('left', 'right'))
    assert tuple(Jl.shape) == (21, 778) and tuple(Jr_.shape) == (21, 778)
    sample = torch.empty(2, B, 8, device=dev)
    roots = torch.empty(2, B, 2, 3, device=dev)
    mrrpe = torch.empty(B, 3, device=dev)
    cdev = torch.empty(B, device=dev)
    pj = torch.empty(2, B, 2, 21, device=dev) if per_element else None
    pv = torch.empty(2, B, 2, 778, device=dev) if per_element else None
    ptrs = (ctypes.c_void_p * 6)(tens[0].data_ptr(), tens[1].data_ptr(), Jl.data_ptr(), tens[2].data_ptr(), tens[3].data_ptr(), Jr_.data_ptr())
    call('rih_eval_metrics', ptrs, B, float(contact_dist), sample.data_ptr(), pj.data_ptr() if per_element else None,
         pv.data_ptr() if per_element else None, roots.data_ptr(), mrrpe.data_ptr(), cdev.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    out = {'sample': sample, 'roots': roots, 'mrrpe': mrrpe, 'cdev': cdev,
           'pajoints_loss': {'left': sample[0, :, 4], 'right': sample[1, :, 4]}, 'paverts_loss': {'left': sample[0, :, 5], 'right': sample[1, :, 5]},
           'double_pa_joint': sample[1, :, 6], 'double_pa_mesh': sample[1, :, 7],
           # the zero-padded left half contributes no error: mean over 2N points = half the right hand's root-relative mean (:549-551)
           'double_joint': sample[1, :, 0] * 0.5, 'double_mesh': sample[1, :, 1] * 0.5}
    if per_element:
        out.update({'orijoint_loss': {'left': pj[0, :, 0], 'right': pj[1, :, 0]}, 'joints_loss': {'left': pj[0, :, 1], 'right': pj[1, :, 1]},
                    'orivert_loss': {'left': pv[0, :, 0], 'right': pv[1, :, 0]}, 'verts_loss': {'left': pv[0, :, 1], 'right': pv[1, :, 1]}})
    return out


class EvalMetrics:
    """Accumulates `batch_metrics` over a test set and reduces like apps/eval_interhand.py:441-552 (values in mm except mrrpe / cdev, which
    the reference prints in metres).  Only the [B,8] / [B,3] / [B] per-sample tensors are kept."""

    def __init__(self, J_regressor, contact_dist=3e-3):
        self.J_regressor, self.contact_dist = J_regressor, contact_dist
        self.sample, self.mrrpe, self.cdev = [], [], []

    def update(self, pred_left, pred_right, gt_left, gt_right):
        m = batch_metrics(self.J_regressor, pred_left, pred_right, gt_left, gt_right, per_element=False, contact_dist=self.contact_dist)
        self.sample.append(m['sample']); self.mrrpe.append(m['mrrpe']); self.cdev.append(m['cdev'])
        return m

    def summary(self, mask=None):
        """mask: optional boolean [N] selection (the reference reports IoU-bucketed subsets with it)."""
        s = torch.cat(self.sample, 1).double()
        mr, cd = torch.cat(self.mrrpe, 0).double(), torch.cat(self.cdev, 0).double()
        if mask is not None:
            mask = torch.as_tensor(mask, device=s.device)
            s, mr, cd = s[:, mask], mr[mask], cd[mask]
        mean = (s.mean(1) * 1000).cpu()
        out = {}
        for h, side in enumerate(('left', 'right')):
            for k, name in enumerate(SAMPLE_FIELDS[:6]):
                out['%s_%s' % (name, side)] = float(mean[h, k])
        for name in SAMPLE_FIELDS[:6]:
            out[name] = (out[name + '_left'] + out[name + '_right']) / 2
        out.update({'double_pa_mpjpe': float(mean[1, 6]), 'double_pa_mpvpe': float(mean[1, 7]),
                    'double_mpjpe': float(mean[1, 0]) / 2, 'double_mpvpe': float(mean[1, 1]) / 2, 'mrrpe': float(mr.mean())})
        ok = ~torch.isnan(cd)
        out['cdev'] = float(cd[ok].sum() / ok.double().sum())
        return out
