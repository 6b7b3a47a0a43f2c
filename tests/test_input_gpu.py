"""GPU input pipeline (csrc/augment.cu rih_augment_u8, csrc/nhwc_ops.cu rih_preprocess_u8; renderih_b200/input.py) against the CPU oracle
(oracle/augment_ref.py, pinned bit-exactly to the reference's handDataset.process_data and to cv2.warpAffine) -- bit-exact for the images."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_augment_matches_reference_golden_bit_exact():
    from oracle import fixtures
    from renderih_b200 import input as I
    gold = torch.load(os.path.join(GOLD, 'augment_synth.pt'), weights_only=False)
    S = gold['samples']
    frames, dicts = fixtures.make_augment_case(3)
    A = np.stack([I.get_affine_mat(s['theta'], s['scale'], s['u'], s['v'], 256, 256) for s in S])
    net, ori, u8 = I.augment_u8(torch.from_numpy(np.stack(frames)).cuda(), A, gain=np.stack([s['a'] for s in S]), offset=[s['b'] for s in S],
                                flip=[s['flip'] for s in S], return_ori=True, return_u8=True)
    for i, s in enumerate(S):
        assert torch.equal(u8[i].cpu(), s['final_u8_bgr']), i
        assert torch.equal(ori[i].cpu(), s['final_u8_bgr'].permute(2, 0, 1).float() / 255)
    assert torch.equal(net[0].cpu(), S[0]['imgTensor'])
    hd = {s: {k: torch.from_numpy(np.stack([d[s][k] for d in dicts])).cuda() for k in dicts[0][s]} for s in ('left', 'right')}
    lab = I.prepare_labels(hd, [s['theta'] for s in S], A, [s['flip'] for s in S], bone_length=gold['meta']['bone_length'])
    for k in lab:
        ref = torch.stack([s['labels'][k] for s in S])
        assert float((lab[k].cpu() - ref).abs().max()) <= 2e-6 * float(ref.abs().max()), k


@pytest.mark.parametrize('B,size', [(64, 256), (5, 96)])
def test_augment_matches_oracle_sweep(B, size):
    """Random maps over (and beyond) the loader's ranges, with / without noise and flip; bit-exact uint8 and float outputs."""
    from oracle import augment_ref as ar
    from renderih_b200 import input as I
    rng = np.random.RandomState(B)
    frames = rng.randint(0, 256, (B, size, size, 3)).astype(np.uint8)
    frames[::2] = (frames[::2].astype(np.float32) * 0.25 + np.linspace(0, 190, size, dtype=np.float32)[None, None, :, None]).astype(np.uint8)
    theta, sc = rng.uniform(-90, 90, B), rng.uniform(0.75, 1.25, B)
    u, v = rng.uniform(-10, 10, B), rng.uniform(-10, 10, B)
    theta[0], sc[0], u[0], v[0] = 0, 1, 0, 0
    theta[1], sc[1], u[1], v[1] = 180, 0.3, size, -size
    flip = rng.rand(B) > 0.5
    a, b = rng.uniform(0.7, 1.3, (B, 3)), 255 * 0.05 * (2 * rng.rand(B) - 1)
    A = np.stack([I.get_affine_mat(theta[i], sc[i], u[i], v[i], size, size) for i in range(B)])
    for noise in (True, False):
        net, u8 = I.augment_u8(torch.from_numpy(frames).cuda(), A, gain=a if noise else None, offset=b if noise else None, flip=flip, return_u8=True)
        for i in range(B):
            img, _, ref_net, M = ar.process_image(frames[i], theta[i], sc[i], u[i], v[i], a[i] if noise else None, b[i], bool(flip[i]))
            assert np.array_equal(M, A[i])
            final = img[:, ::-1] if flip[i] else img
            assert np.array_equal(u8[i].cpu().numpy(), final), (i, noise)
            assert np.array_equal(net[i].cpu().numpy(), ref_net), (i, noise)
    # identity map + no noise + no flip == the plain preprocess kernel
    ident = np.stack([np.eye(3, dtype=np.float32)] * B)
    assert torch.equal(I.augment_u8(torch.from_numpy(frames).cuda(), ident), I.preprocess_u8(torch.from_numpy(frames).cuda()))


def test_augment_rejects_cpu_input():
    from renderih_b200 import input as I
    with pytest.raises(RuntimeError):
        I.augment_u8(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), np.eye(3, dtype=np.float32)[None])
