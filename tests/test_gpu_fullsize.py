"""Parity at BASELINE.json's full sizes (batch 256 per GPU), where the CPU oracle would take minutes: size-independent
properties of the domain instead of element-wise comparison --
  * softmax rows sum to one (V = 1  =>  O = 1), key-permutation invariance, sum_k dV[k] = sum_q dO[q];
  * the tcgen05 attention kernels against the independent mma.sync kernels of the same library;
  * masking: ids_restore is a permutation, exactly len_keep zeros per row, ids_keep = the len_keep smallest noises;
  * the GEMM against an fp32 cuBLAS product of the same bf16 operands;
  * a batch-256 MAE loss equals the mean of the losses of its two halves (a checksum of checksums: every sample masks
    the same number of patches), and one optimizer step leaves every parameter finite."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _attn(qkv, B, S, H, hd, dout=None):
    from headct_foundation_b200._cabi import call, stream_ptr
    dev = qkv.device
    out = torch.empty(B, S, H * hd, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, S, device=dev)
    call("hct_attention_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, S, H, hd, stream_ptr(dev))
    if dout is None:
        return out, lse
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(B, H, S, device=dev)
    call("hct_attention_bwd", qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
         delta.data_ptr(), B, S, H, hd, stream_ptr(dev))
    return out, lse, dqkv


@pytest.mark.parametrize("B,S,H,hd", [(256, 513, 16, 48), (256, 129, 12, 64)])
def test_attention_full_size_properties(cuda, B, S, H, hd):
    from headct_foundation_b200._cabi import lib
    D = H * hd
    g = torch.Generator(device="cuda").manual_seed(7)
    qkv = torch.randn(B, S, 3, D, device=cuda, generator=g).bfloat16()
    # (1) V = 1  =>  every output element is 1 (rows of P sum to one); lse finite
    q1 = qkv.clone()
    q1[:, :, 2] = 1.0
    out, lse = _attn(q1.view(B, S, 3 * D), B, S, H, hd)
    assert torch.isfinite(lse).all()
    assert (out.float() - 1.0).abs().max().item() < 1.6e-2          # P is rounded to bf16 before P V
    # (2) permuting the keys (K and V rows together) changes nothing but the summation order
    perm = torch.randperm(S, device=cuda, generator=g)
    qp = qkv.clone()
    qp[:, :, 1:] = qkv[:, perm][:, :, 1:]
    o0, l0 = _attn(qkv.view(B, S, 3 * D), B, S, H, hd)
    o1, l1 = _attn(qp.view(B, S, 3 * D), B, S, H, hd)
    assert _rel(o1.float(), o0.float()) < 6e-3
    assert (l1 - l0).abs().max().item() < 1e-4
    # (3) backward: sum over keys of dV equals sum over queries of dO; tcgen05 kernels vs the mma.sync kernels
    dout = torch.randn(B, S, D, device=cuda, generator=g).bfloat16()
    _, _, dqkv = _attn(qkv.view(B, S, 3 * D), B, S, H, hd, dout)
    dv = dqkv.view(B, S, 3, D)[:, :, 2].float()
    assert _rel(dv.sum(1), dout.float().sum(1)) < 5e-3
    lib().hct_attention_set_tcgen05(0)
    try:
        o_ref, l_ref, d_ref = _attn(qkv.view(B, S, 3 * D), B, S, H, hd, dout)
    finally:
        lib().hct_attention_set_tcgen05(2)
    assert _rel(o0.float(), o_ref.float()) < 6e-3 and (l0 - l_ref).abs().max().item() < 1e-4
    for i, name in enumerate("qkv"):
        a, r = dqkv.view(B, S, 3, D)[:, :, i].float(), d_ref.view(B, S, 3, D)[:, :, i].float()
        assert _rel(a, r) < 1.2e-2, name


def test_masking_full_size_properties(cuda):
    from headct_foundation_b200 import functional as HF
    N, L, keep = 256 * 8, 512, 128
    g = torch.Generator(device="cuda").manual_seed(11)
    noise = torch.rand(N, L, device=cuda, generator=g)
    noise[:, 100] = noise[:, 7]                                       # a forced tie in every row
    ids_restore, ids_keep, mask = HF.mask_indices(noise, keep)
    assert torch.equal(torch.sort(ids_restore, dim=1).values, torch.arange(L, device=cuda).expand(N, L))
    assert torch.equal(mask.sum(1), torch.full((N,), float(L - keep), device=cuda))
    assert torch.equal(mask, (ids_restore >= keep).float())
    # ids_keep = positions of the `keep` smallest noises in stable ascending order = inverse permutation prefix
    ids_shuffle = torch.argsort(noise, dim=1, stable=True)
    assert torch.equal(ids_keep, ids_shuffle[:, :keep])
    assert torch.equal(torch.gather(ids_restore, 1, ids_keep), torch.arange(keep, device=cuda).expand(N, keep))


@pytest.mark.parametrize("M,N,K", [(131328, 768, 768), (33024, 3072, 768), (131328, 768, 3072)])
def test_gemm_full_size_against_fp32_product(cuda, M, N, K):
    from headct_foundation_b200 import functional as HF
    g = torch.Generator(device="cuda").manual_seed(M % 1000 + N)
    A = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    B = (torch.randn(N, K, device=cuda, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=cuda, generator=g)
    out = torch.empty(M, N, device=cuda, dtype=torch.float32)
    HF.gemm(A, B, M=M, N=N, K=K, lda=K, ldb=K, out=out, ldo=N, epi=HF.EPI_F32, bias=bias)
    ref = A.float() @ B.float().t() + bias
    assert _rel(out, ref) < 1e-5
    # linearity in A: C(A1 + A2) = C(A1) + C(A2) - bias  (bf16 inputs chosen so that A1 + A2 is exact)
    A2 = (A.float() * 0.5).bfloat16()
    out2 = torch.empty_like(out)
    HF.gemm(A2, B, M=M, N=N, K=K, lda=K, ldb=K, out=out2, ldo=N, epi=HF.EPI_F32, bias=bias)
    assert _rel(2.0 * (out2 - bias), out - bias) < 1e-5


def test_mae_batch256_loss_is_mean_of_half_batches_and_step_is_finite(cuda):
    import headct_foundation_b200 as H
    from headct_foundation_b200.configs import MAE_HEADCT
    from headct_foundation_b200.optim import FusedAdamW
    torch.manual_seed(3)
    model = H.MaskedAutoencoderViT(**MAE_HEADCT).to(cuda).train()
    B = 256
    x = torch.rand(B, 3, 96, 96, 96, device=cuda)
    noise = torch.rand(B, 512, device=cuda)

    def loss_of(xs, ns):
        model.noise_override = ns
        try:
            with torch.no_grad():
                return model(xs)[0].item()
        finally:
            model.noise_override = None

    whole = loss_of(x, noise)
    halves = 0.5 * (loss_of(x[:128], noise[:128]) + loss_of(x[128:], noise[128:]))
    assert abs(whole - halves) <= 2e-5 * abs(whole)
    opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1.5e-4, betas=(0.9, 0.95), eps=1e-8,
                     weight_decay=0.05, clip_grad=3.0)
    model.noise_override = noise
    loss, _, _ = model(x)
    loss.backward()
    model.noise_override = None
    assert abs(loss.item() - whole) <= 1e-3 * abs(whole)              # training forward (saves GELU') = inference forward
    opt.step()
    for n_, p in model.named_parameters():
        assert torch.isfinite(p).all(), n_
        if p.requires_grad:
            assert p.grad is not None and torch.isfinite(p.grad).all(), n_
