"""Debug aid: ResNet bottleneck with a stride-2 shortcut, every GEMM mode / switch combination against a plain torch fp32 reference of the block."""
import itertools, os, sys, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from renderih_b200 import ops
from renderih_b200._lib import call
from renderih_b200.model import ResNetSimple

DEV = 'cuda:0'
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))   # Frobenius: robust to single ReLU-gate flips


def torch_block(blk, x_rows, N, H):
    x = x_rows.reshape(N, H, H, -1).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    def cb(t, conv, bn, relu=True):
        t = F.conv2d(t, conv.weight, None, stride=conv.stride, padding=conv.padding)
        t = F.batch_norm(t, None, None, bn.weight, bn.bias, True, 0.1, bn.eps)
        return F.relu(t) if relu else t
    o = cb(x, blk.conv1, blk.bn1)
    o = cb(o, blk.conv2, blk.bn2)
    o = cb(o, blk.conv3, blk.bn3, relu=False)
    idn = cb(x, blk.downsample[0], blk.downsample[1], relu=False)
    y = F.relu(o + idn)
    return x, y


def main():
    torch.manual_seed(0)
    enc = ResNetSimple('resnet50', aux_heads=False).to(DEV).train()
    blk = enc.resnet.layer2[0]
    N, H = 4, 64
    x0 = torch.randn(N * H * H, 256, device=DEV) * 0.5
    g = torch.Generator(device='cpu').manual_seed(1)
    wgt = torch.randn(N * 32 * 32, 512, generator=g).to(DEV)
    for p in blk.parameters():
        p.grad = None
    xr, yr = torch_block(blk, x0, N, H)
    yr_rows = yr.permute(0, 2, 3, 1).reshape(-1, 512)
    (yr_rows * wgt).sum().backward()
    ref = (yr_rows.detach(), xr.grad.permute(0, 2, 3, 1).reshape(-1, 256).clone(), {k: p.grad.clone() for k, p in blk.named_parameters()})
    for mode, direct, alias in itertools.product(('simt', 'tf32x3'), (1, 0), (1, 0)):
        for p in blk.parameters():
            p.grad = None
        call('rih_set_s2_direct', direct)
        x = x0.clone().requires_grad_(True)
        ops.set_gemm_mode(mode, mode)
        try:
            y, Ho = enc._bottleneck(blk, x, N, H) if alias else _no_alias(enc, blk, x, N, H)
            (y * wgt).sum().backward()
        finally:
            ops.set_gemm_mode('simt', 'simt')
            call('rih_set_s2_direct', 1)
        torch.cuda.synchronize()
        line = ['%-7s direct=%d alias=%d' % (mode, direct, alias), 'y %.1e' % rel(y.detach(), ref[0]), 'dx %.1e' % rel(x.grad, ref[1])]
        for k, p in blk.named_parameters():
            line.append('%s %.1e' % (k.replace('.weight', '.w').replace('.bias', '.b').replace('downsample', 'ds'), rel(p.grad, ref[2][k])))
        print('  '.join(line), flush=True)


def _no_alias(enc, blk, x, N, H):
    from renderih_b200.model import _conv_bn
    tr = True
    out, _ = _conv_bn(x, blk.conv1, blk.bn1, N, H, H, tr)
    out, Ho = _conv_bn(out, blk.conv2, blk.bn2, N, H, H, tr)
    identity, _ = _conv_bn(x, blk.downsample[0], blk.downsample[1], N, H, H, tr, relu=False)
    out, _ = _conv_bn(out, blk.conv3, blk.bn3, N, Ho, Ho, tr, relu=True, res=identity)
    return out, Ho


if __name__ == '__main__':
    main()
