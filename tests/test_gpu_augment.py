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


def test_crop_resize_area_matches_torch_area_interpolation(cuda):
    from headct_foundation_b200 import functional as HF
    from oracle import headct_oracle as O
    g = torch.Generator().manual_seed(9)
    src = torch.rand(3, 2, 40, 36, 44, generator=g)
    boxes = torch.tensor([[0, 0, 0, 0, 40, 36, 44],        # whole volume, downsample
                          [1, 5, 3, 7, 24, 24, 24],        # interior crop, same size as the target
                          [2, -6, -4, 10, 30, 50, 40],     # runs over the border on three sides: zero padding
                          [0, 10, 10, 10, 13, 17, 11],     # small crop, upsampled (area windows of 1-2 voxels)
                          [1, 38, 0, 0, 9, 36, 44]], dtype=torch.int32)
    for dt in (torch.float32, torch.float16):
        s = src.to(dt)
        got = HF.crop_resize_area(s.to(cuda), boxes, (24, 24, 24)).cpu()
        want = O.crop_resize_area(s, boxes, (24, 24, 24))
        assert got.shape == (5, 2, 24, 24, 24)
        assert (got - want).abs().max().item() < 2e-6


def test_crop_resize_area_row_staged_kernel(cuda):
    """Source rows of a multiple of 8 elements take the row-staged kernel (one 16-byte load per lane and source row):
    same result as the per-voxel gather and as torch's area interpolation, flips and offsets included."""
    from headct_foundation_b200 import functional as HF
    from headct_foundation_b200._cabi import lib
    from oracle import headct_oracle as O
    g = torch.Generator().manual_seed(10)
    src = torch.rand(3, 2, 40, 36, 48, generator=g)
    boxes = torch.tensor([[0, 0, 0, 0, 40, 36, 48], [1, 5, 3, 7, 24, 24, 24], [2, -6, -4, 10, 30, 50, 46],
                          [0, 10, 10, 13, 13, 17, 11], [1, 38, 0, 3, 9, 36, 44], [2, 1, 2, -5, 33, 30, 60]], dtype=torch.int32)
    flips = torch.tensor([0, 7, 1, 4, 2, 5], dtype=torch.int32)
    offs = torch.tensor([0.0, 0.05, -0.03, 0.1, 0.0, -0.08])
    for dt in (torch.float32, torch.float16):
        s = src.to(dt)
        want = O.crop_resize_area(s, boxes, (24, 20, 28))
        for ax in range(3):
            sel = (flips >> ax) & 1
            want = torch.where(sel.bool().view(-1, 1, 1, 1, 1), want.flip(2 + ax), want)
        want = want + offs.view(-1, 1, 1, 1, 1)
        got = HF.crop_resize_area(s.to(cuda), boxes, (24, 20, 28), flips, offs).cpu()
        lib().hct_crop_resize_set_rows(0)
        try:
            old = HF.crop_resize_area(s.to(cuda), boxes, (24, 20, 28), flips, offs).cpu()
        finally:
            lib().hct_crop_resize_set_rows(1)
        assert (old - want).abs().max().item() < 2e-6
        assert (got - want).abs().max().item() < 2e-6 and (got - old).abs().max().item() < 2e-6
    big = torch.rand(2, 1, 64, 64, 224, generator=g).half()                     # the cached-volume row length
    bx = torch.tensor([[0, 3, 5, 17, 60, 50, 190], [1, -4, 0, 100, 64, 64, 160]], dtype=torch.int32)
    got = HF.crop_resize_area(big.to(cuda), bx, (32, 32, 96)).cpu()
    assert (got - O.crop_resize_area(big, bx, (32, 32, 96))).abs().max().item() < 2e-6


def test_adjust_contrast_matches_oracle(cuda):
    from headct_foundation_b200 import functional as HF
    from oracle import headct_oracle as O
    g = torch.Generator().manual_seed(10)
    vol = torch.rand(5, 3, 16, 16, 16, generator=g) * 1.4 - 0.2
    gamma = torch.tensor([0.2, 0.0, 1.0, 0.55, 0.0])
    got = HF.adjust_contrast_(vol.clone().to(cuda), gamma).cpu()
    want = O.adjust_contrast(vol, gamma)
    assert torch.equal(got[1], vol[1]) and torch.equal(got[4], vol[4])
    assert (got - want).abs().max().item() < 2e-6


def test_dino_multicrop_end_to_end(cuda):
    """DataAugmentationDINO3D on a batch: every crop equals the oracle chain (pad/crop -> area resize -> flip -> shift ->
    smooth / contrast) run with the same draws; a volume smaller than the 224 canvas exercises the zero padding."""
    from headct_foundation_b200.data.transforms import DataAugmentationDINO3D
    from oracle import headct_oracle as O
    aug = DataAugmentationDINO3D((32, 32, 32), global_crops_size=112, local_crops_size=64, local_crops_number=2, seed=3)
    B = 6
    vol = torch.rand(B, 2, 200, 180, 230).half()
    for trial in range(3):                                   # a few draw sets so that smoothing and contrast both occur
        draws = aug.randomize(B, vol.shape[2:])
        if bool((draws["sigma"][:, 0] > 0).any()) and bool((draws["gamma"] > 0).any()):
            break
    crops = aug.apply(vol.to(cuda), draws)
    assert len(crops) == 4 and all(c.shape == (B, 2, 32, 32, 32) and c.dtype == torch.float32 for c in crops)
    bx = draws["boxes"].view(4, B, 7)
    assert int(bx[:2, :, 4:].min()) >= 112 and int(bx[:2, :, 4:].max()) <= 224        # global crop sizes
    assert int(bx[2:, :, 4:].min()) >= 64 and int(bx[2:, :, 4:].max()) <= 112         # local crop sizes
    for cidx in range(4):
        want = O.crop_resize_area(vol, bx[cidx], (32, 32, 32))
        want = O.flip_shift(want, draws["flips"].view(4, B)[cidx], draws["offsets"].view(4, B)[cidx])
        if cidx == 0:
            want = O.gaussian_smooth(want, draws["sigma"])
        if cidx == 1:
            want = O.adjust_contrast(want, draws["gamma"])
        assert (crops[cidx].cpu() - want).abs().max().item() < 5e-6, cidx
    out = aug(vol.to(cuda))
    assert len(out) == 4 and out[0].is_cuda
