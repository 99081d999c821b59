"""SURVEY 8(f) rank 3 (MAE / ViT chain): on-GPU RandFlip x 3 + RandShiftIntensity + RandGaussianSmooth of
mae3d_transforms (src/data/transforms.py:195-236) against the CPU restatement in the oracle.  Flips and the shift are
exact data movement + one fp32 add: bit-exact.  The Gaussian is a separable fp32 convolution: 1e-6."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("shape", [(9, 3, 24, 20, 16), (2, 1, 8, 8, 8), (3, 3, 96, 96, 96), (8, 2, 5, 7, 24)])
def test_flip_shift_bit_exact(cuda, dtype, shape):
    from headct_foundation_b200 import functional as HF
    from oracle import headct_oracle as O
    g = torch.Generator().manual_seed(shape[0])
    vol = torch.rand(*shape, generator=g).to(dtype)
    flips = torch.arange(shape[0], dtype=torch.uint8) % 8            # every combination of the three axes
    offs = (torch.rand(shape[0], generator=g) - 0.5) * 0.2
    offs[0] = 0.0
    got = HF.flip_shift(vol.to(cuda), flips, offs)
    want = O.flip_shift(vol, flips, offs)
    assert got.dtype == torch.float32 and torch.equal(got.cpu(), want)
    assert torch.equal(HF.flip_shift(vol.to(cuda), None, None).cpu(), vol.float())


def test_gaussian_smooth_matches_oracle(cuda):
    from headct_foundation_b200 import functional as HF
    from headct_foundation_b200.data.transforms import gaussian_taps
    from oracle import headct_oracle as O
    g = torch.Generator().manual_seed(4)
    vol = torch.rand(4, 3, 20, 24, 16, generator=g)
    sig = torch.tensor([[0.5, 0.75, 1.0], [0.0, 0.0, 0.0], [1.0, 1.0, 0.5], [0.9, 0.6, 0.7]])
    assert torch.equal(gaussian_taps(0.8), O.gaussian_kernel_1d(0.8))
    radius = 4
    on = torch.tensor([0, 2, 3], dtype=torch.int32)          # sample 1 is not smoothed and must stay untouched
    taps = [torch.stack([gaussian_taps(float(sig[b, k]), radius=radius) for b in on.tolist()]) for k in range(3)]
    dev_vol = vol.to(cuda)
    got = HF.gaussian_smooth(dev_vol, taps, on).cpu()
    want = O.gaussian_smooth(vol, sig)
    assert (got - want).abs().max().item() < 1e-6
    assert torch.equal(got[1], vol[1])


def test_mae3d_train_augment_end_to_end(cuda):
    from headct_foundation_b200.data.transforms import MAE3DTrainAugment
    from oracle import headct_oracle as O
    aug = MAE3DTrainAugment(prob_flip=0.5, prob_shift=0.5, prob_smooth=0.5, seed=7)
    vol = torch.rand(8, 3, 24, 24, 24).half()
    flips, offs, sig = aug.randomize(8)
    assert flips.dtype == torch.uint8 and int(flips.max()) < 8 and bool((sig.sum(1) > 0).any()) and bool((sig.sum(1) == 0).any())
    got = aug.apply(vol.to(cuda), flips, offs, sig).cpu()
    want = O.gaussian_smooth(O.flip_shift(vol, flips, offs), sig)
    assert (got - want).abs().max().item() < 1e-6
    out = aug({"image": vol.to(cuda)})["image"]
    assert out.shape == vol.shape and out.dtype == torch.float32 and out.is_cuda
