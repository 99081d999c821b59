"""GPU parity tests for the individual C-ABI kernels against fp32 torch restatements of the same op."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def HF():
    from headct_foundation_b200 import functional
    return functional


# ------------------------------------------------------------------ GEMM
GEMM_SHAPES = [(128, 256, 64), (256, 768, 768), (1000, 2304, 768), (130, 96, 192), (513, 5184, 768), (77, 64, 96),
               (4096, 3072, 768), (384, 768, 3072)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_kk_bf16(cuda, HF, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    B = torch.randn(N, K, device=cuda, generator=g).bfloat16()
    bias = torch.randn(N, device=cuda, generator=g)
    out = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
    HF.gemm(A, B, M=M, N=N, K=K, lda=K, ldb=K, out=out, ldo=N, epi=HF.EPI_BF16, bias=bias)
    ref = A.float() @ B.float().t() + bias
    assert _rel(out.float(), ref) < 6e-3


@pytest.mark.parametrize("M,N,K", [(256, 768, 768), (1000, 768, 2304), (130, 192, 96), (513, 768, 5184)])
def test_gemm_dgrad_b_mn_major(cuda, HF, M, N, K):
    """dX[M,N] = dY[M,K] @ W[K,N]  with W given as stored ([K,N] row-major == MN-major B)."""
    g = torch.Generator(device="cuda").manual_seed(N + K)
    dY = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    W = torch.randn(K, N, device=cuda, generator=g).bfloat16()
    out = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
    HF.gemm(dY, W, M=M, N=N, K=K, lda=K, ldb=N, b_mn=True, out=out, ldo=N, epi=HF.EPI_BF16)
    ref = dY.float() @ W.float()
    assert _rel(out.float(), ref) < 6e-3


@pytest.mark.parametrize("T,N,K", [(512, 768, 768), (1000, 2304, 768), (4104, 768, 3072), (130, 96, 192),
                                   (33024, 768, 768), (1026, 5184, 768)])
def test_gemm_wgrad_mn_mn_splitk(cuda, HF, T, N, K):
    """dW[N,K] = dY[T,N]^T @ X[T,K]  (both MN-major, split-K with fp32 red.add)."""
    g = torch.Generator(device="cuda").manual_seed(T + N)
    dY = torch.randn(T, N, device=cuda, generator=g).bfloat16()
    X = torch.randn(T, K, device=cuda, generator=g).bfloat16()
    out = HF.linear_wgrad(dY, X)
    ref = dY.float().t() @ X.float()
    assert _rel(out, ref) < 2e-3


def test_gemm_epilogues(cuda, HF):
    M, N, K = 640, 768, 384
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    B = (torch.randn(N, K, device=cuda, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=cuda, generator=g)
    acc = A.float() @ B.float().t()
    # GELU (+ pre-activation side output)
    out = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
    pre = torch.empty_like(out)
    HF.gemm(A, B, M=M, N=N, K=K, lda=K, ldb=K, out=out, ldo=N, epi=HF.EPI_GELU_BF16, bias=bias, out2=pre, ldo2=N)
    assert _rel(pre.float(), acc + bias) < 6e-3
    assert _rel(out.float(), torch.nn.functional.gelu(acc + bias)) < 8e-3
    # residual fp32, in place
    res = torch.randn(M, N, device=cuda, generator=g)
    want = res + acc + bias
    HF.gemm(A, B, M=M, N=N, K=K, lda=K, ldb=K, out=res, ldo=N, epi=HF.EPI_RES_F32, bias=bias, res=res, ldres=N)
    assert _rel(res, want) < 1e-5 + 2e-3
    # dGELU
    aux = torch.randn(M, N, device=cuda, generator=g).bfloat16()
    HF.gemm(A, B, M=M, N=N, K=K, lda=K, ldb=K, out=out, ldo=N, epi=HF.EPI_DGELU_BF16, aux=aux, ldaux=N)
    a = aux.float().requires_grad_(True)
    torch.nn.functional.gelu(a).backward(acc)
    assert _rel(out.float(), a.grad) < 8e-3
    # GELU + its derivative as the side output (training forward), then the multiplying dgrad epilogue + column sums
    drv = torch.empty_like(out)
    HF.gemm(A, B, M=M, N=N, K=K, lda=K, ldb=K, out=out, ldo=N, epi=HF.EPI_GELU_DERIV_BF16, bias=bias, out2=drv, ldo2=N)
    t = (acc + bias).requires_grad_(True)
    gl = torch.nn.functional.gelu(t)
    gl.backward(torch.ones_like(gl))
    assert _rel(out.float(), gl.detach()) < 8e-3
    assert _rel(drv.float(), t.grad) < 8e-3
    assert (out.float() - gl.detach()).abs().max() < 2e-2 and (drv.float() - t.grad).abs().max() < 1e-2
    cs = torch.zeros(N, device=cuda)
    HF.gemm(A, B, M=M, N=N, K=K, lda=K, ldb=K, out=out, ldo=N, epi=HF.EPI_MUL_BF16, aux=aux, ldaux=N, colsum=cs)
    assert _rel(out.float(), acc * aux.float()) < 8e-3
    assert _rel(cs, out.float().sum(0)) < 1e-4
    # ragged edges (M, N not multiples of the 32-wide epilogue units) take the general epilogue path
    Mr, Nr = 77, 40
    o_r, d_r = torch.empty(Mr, Nr, device=cuda, dtype=torch.bfloat16), torch.empty(Mr, Nr, device=cuda, dtype=torch.bfloat16)
    HF.gemm(A, B, M=Mr, N=Nr, K=K, lda=K, ldb=K, out=o_r, ldo=Nr, epi=HF.EPI_GELU_DERIV_BF16, bias=bias, out2=d_r, ldo2=Nr)
    assert _rel(o_r.float(), gl.detach()[:Mr, :Nr]) < 8e-3 and _rel(d_r.float(), t.grad[:Mr, :Nr]) < 8e-3
    # fp32 store + position table with row remap (3 prefix rows per group of 64)
    pos = torch.randn(64, N, device=cuda, generator=g)
    groups = M // 64
    o32 = torch.full((groups, 67, N), 7.0, device=cuda)
    HF.gemm(A, B, M=M, N=N, K=K, lda=K, ldb=K, out=o32, ldo=N, epi=HF.EPI_POS_F32, bias=bias, pos=pos, ldpos=N,
            pos_period=64, rows_in=64, rows_out=67, row_off=3)
    want = (acc + bias).view(groups, 64, N) + pos
    assert _rel(o32[:, 3:], want) < 2e-3
    assert torch.all(o32[:, :3] == 7.0)
    # explicit per-row position index
    idx = torch.randint(0, 64, (M,), device=cuda, dtype=torch.int32)
    o2 = torch.empty(M, N, device=cuda)
    HF.gemm(A, B, M=M, N=N, K=K, lda=K, ldb=K, out=o2, ldo=N, epi=HF.EPI_POS_F32, bias=bias, pos=pos, ldpos=N, pos_idx=idx)
    assert _rel(o2, acc + bias + pos[idx.long()]) < 2e-3


def test_gemm_fused_colsum(cuda, HF):
    M, N, K = 1000, 768, 256
    g = torch.Generator(device="cuda").manual_seed(9)
    A = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    B = (torch.randn(N, K, device=cuda, generator=g) / math.sqrt(K)).bfloat16()
    out = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
    cs = torch.zeros(N, device=cuda)
    HF.gemm(A, B, M=M, N=N, K=K, lda=K, ldb=K, out=out, ldo=N, epi=HF.EPI_BF16, colsum=cs)
    assert _rel(cs, out.float().sum(0)) < 1e-5


# ------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("rows,dim,eps", [(1000, 768, 1e-5), (333, 96, 1e-6), (64, 192, 1e-5), (50, 1024, 1e-5),
                                          (5000, 768, 1e-5)])
def test_layernorm_fwd_bwd(cuda, HF, rows, dim, eps):
    g = torch.Generator(device="cuda").manual_seed(rows)
    x = torch.randn(rows, dim, device=cuda, generator=g) * 2 + 0.5
    w = torch.randn(dim, device=cuda, generator=g)
    b = torch.randn(dim, device=cuda, generator=g)
    y, mean, rstd = HF.layernorm_fwd(x, w, b, eps, False, True)
    xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (dim,), wr, br, eps)
    assert _rel(y, yr.detach()) < 1e-5
    y16, _, _ = HF.layernorm_fwd(x, w, b, eps, True, False)
    assert _rel(y16.float(), yr.detach()) < 5e-3
    dy = torch.randn(rows, dim, device=cuda, generator=g)
    dres = torch.randn(rows, dim, device=cuda, generator=g)
    yr.backward(dy)
    dx, dx16, dg, db = HF.layernorm_bwd(dy, x, w, mean, rstd, dres, True)
    assert _rel(dx, xr.grad + dres) < 1e-5
    assert _rel(dx16.float(), xr.grad + dres) < 5e-3
    assert _rel(dg, wr.grad) < 1e-4 and _rel(db, br.grad) < 1e-4
    _, d16, _, _, dsum = HF.layernorm_bwd(dy, x, w, mean, rstd, dres, True, want_colsum=True)
    assert _rel(dsum, d16.float().sum(0)) < 1e-5
    dx_b, _, _, _ = HF.layernorm_bwd(dy.bfloat16(), x, w, mean, rstd, None, False)
    xr.grad = None
    torch.nn.functional.layer_norm(xr, (dim,), wr, br, eps).backward(dy.bfloat16().float())
    assert _rel(dx_b, xr.grad) < 1e-5


# ------------------------------------------------------------------ attention
@pytest.mark.parametrize("mode", [2, 3, 1, 0])
@pytest.mark.parametrize("B,S,H,hd", [(2, 129, 12, 64), (2, 513, 16, 48), (1, 517, 12, 64), (3, 65, 2, 48),
                                      (2, 260, 2, 48), (1, 136, 3, 64),      # 4 / 8 rows behind the last full tile
                                      (3, 17, 3, 64), (2, 64, 4, 32), (1, 1, 2, 64),
                                      # ragged tails: 38 / 8 / 44 / exactly 16 valid keys in the last 64-key block,
                                      # query tiles with 1-4 live warps, and exact multiples of the tile sizes
                                      (2, 230, 4, 48), (2, 200, 3, 64), (1, 300, 2, 48), (2, 144, 2, 64),
                                      (1, 128, 2, 64), (2, 256, 2, 48), (1, 385, 1, 64),
                                      # S = 128 k + 1 and S = 64 k + 1: the forward folds the last key into its epilogue and
                                      # lets the last query row ride with the last full tile's CTA (hct_attention_set_tail_key)
                                      (2, 257, 3, 48), (1, 193, 2, 64), (3, 641, 2, 48), (5, 129, 3, 48)])
def test_attention_fwd_bwd(cuda, HF, B, S, H, hd, mode):
    """mode 2 (default): tcgen05 kernels, backward rows behind the last full 128-row tile on the row kernel; 3: tcgen05 for
    every tile; 1: forward tail rows on mma.sync; 0: mma.sync kernels only.  Backward on the two-CTA-per-SM kernels here
    (hct_attention_set_bwd3(0)); the default pipelined persistent backward has its own test below."""
    from headct_foundation_b200._cabi import call, stream_ptr, lib
    lib().hct_attention_set_bwd3(0)
    if mode != 2 and (hd == 32 or S < 64):
        pytest.skip("mode only matters for the tcgen05 shapes")
    lib().hct_attention_set_tcgen05(mode)
    try:
        _attention_case(cuda, B, S, H, hd)
    finally:
        lib().hct_attention_set_tcgen05(2)
        lib().hct_attention_set_bwd3(1)


@pytest.mark.parametrize("B,S,H,hd", [(2, 129, 12, 64), (2, 513, 16, 48), (1, 517, 12, 64), (2, 65, 2, 48), (1, 321, 2, 64),
                                      (1, 200, 3, 64)])
def test_attention_bwd_unmerged_tail_block(cuda, HF, B, S, H, hd):
    """The default backward computes a <= 16-row last block together with block 0 (hct_attention_set_merge_tail(1));
    the older schedule -- the tail block as its own chain step -- stays available for A/B timing and must agree too."""
    from headct_foundation_b200._cabi import lib
    lib().hct_attention_set_bwd3(0)
    lib().hct_attention_set_merge_tail(0)
    try:
        _attention_case(cuda, B, S, H, hd)
    finally:
        lib().hct_attention_set_merge_tail(1)
        lib().hct_attention_set_bwd3(1)


@pytest.mark.parametrize("B,S,H,hd", [(2, 129, 12, 64), (2, 513, 16, 48), (1, 517, 12, 64), (2, 65, 2, 48), (1, 321, 2, 64),
                                      (1, 200, 3, 64), (2, 230, 4, 48), (1, 128, 2, 64)])
def test_attention_bwd_pipelined_dkdv_variant(cuda, HF, B, S, H, hd):
    """The off-by-default dK/dV kernel with 32-query blocks and two S^T / dP^T buffer pairs (hct_attention_set_dkdv32)."""
    from headct_foundation_b200._cabi import lib
    lib().hct_attention_set_bwd3(0)
    lib().hct_attention_set_dkdv32(1)
    try:
        _attention_case(cuda, B, S, H, hd)
    finally:
        lib().hct_attention_set_dkdv32(0)
        lib().hct_attention_set_bwd3(1)


@pytest.mark.parametrize("mode", [2, 3])
@pytest.mark.parametrize("B,S,H,hd", [(2, 129, 12, 64), (2, 513, 16, 48), (1, 517, 12, 64), (3, 65, 2, 48), (2, 260, 2, 48),
                                      (1, 136, 3, 64), (3, 17, 3, 64), (1, 1, 2, 64), (2, 230, 4, 48), (2, 200, 3, 64),
                                      (1, 300, 2, 48), (2, 144, 2, 64), (1, 128, 2, 64), (2, 256, 2, 48), (1, 385, 1, 64),
                                      # many more work items than SMs: every CTA walks several (batch, head, tile) items,
                                      # with 1, 2, 3 and 9 streamed blocks per item (ring / buffer phases wrap around)
                                      (40, 129, 12, 64), (12, 513, 16, 48), (90, 40, 4, 64), (70, 100, 3, 48), (9, 517, 12, 64)])
def test_attention_bwd_pipelined_persistent(cuda, HF, B, S, H, hd, mode):
    """hct_attention_set_bwd3(1): one persistent CTA per SM, three score-buffer pairs in tensor memory, two softmax groups
    (csrc/hct_attention_bwd3.cu) against the fp32 torch reference, with and without the row kernel for the tail rows."""
    from headct_foundation_b200._cabi import lib
    lib().hct_attention_set_bwd3(1)
    lib().hct_attention_set_tcgen05(mode)
    try:
        _attention_case(cuda, B, S, H, hd)
    finally:
        lib().hct_attention_set_tcgen05(2)


@pytest.mark.parametrize("B,S,H,hd", [(2, 129, 12, 64), (2, 513, 16, 48), (1, 517, 12, 64), (3, 65, 2, 48), (2, 260, 2, 48),
                                      (1, 136, 3, 64), (3, 17, 3, 64), (1, 1, 2, 64), (2, 230, 4, 48), (2, 200, 3, 64),
                                      (1, 300, 2, 48), (2, 144, 2, 64), (1, 128, 2, 64), (2, 256, 2, 48), (1, 385, 1, 64),
                                      (1, 96, 2, 48), (2, 640, 2, 64), (1, 1000, 1, 48),
                                      # many more work items than SMs: every CTA walks several (batch, head, tile pair) items
                                      (40, 129, 12, 64), (12, 513, 16, 48), (90, 40, 4, 64), (70, 100, 3, 48), (9, 517, 12, 64)])
def test_attention_fwd_pipelined_persistent(cuda, HF, B, S, H, hd):
    """Forward on the pipelined persistent kernel (hct_attention_set_fwd2(1): two query tiles of a head per CTA, two score
    buffers each, sixteen softmax warps) against the fp32 torch attention; the backward that follows in the case consumes its
    outputs and log-sum-exp."""
    from headct_foundation_b200._cabi import lib
    lib().hct_attention_set_fwd2(1)
    lib().hct_attention_set_tcgen05(3)
    try:
        _attention_case(cuda, B, S, H, hd)
    finally:
        lib().hct_attention_set_tcgen05(2)
        lib().hct_attention_set_fwd2(0)


@pytest.mark.parametrize("B,S,H,hd", [(3, 129, 12, 64), (2, 513, 16, 48), (2, 257, 3, 48), (1, 193, 2, 64), (2, 65, 2, 48)])
def test_attention_fwd_tail_folds_agree_with_own_block_and_tile(cuda, B, S, H, hd):
    """hct_attention_set_tail_key: bit 0 folds the single key behind the last 64-key block into the epilogue, bit 1 lets the
    single query row behind the last full tile ride with that tile's CTA.  Both against the kernel that gives them their own
    block / tile: outputs within the bf16 rounding of the probabilities involved, log-sum-exp to fp32 rounding."""
    from headct_foundation_b200._cabi import call, stream_ptr, lib
    D = H * hd
    g = torch.Generator(device="cuda").manual_seed(S + hd)
    qkv = torch.randn(B, S, 3 * D, device=cuda, generator=g).bfloat16()
    res = {}
    try:
        for fold in (0, 1, 3):
            lib().hct_attention_set_tail_key(fold)
            out = torch.empty(B, S, D, device=cuda, dtype=torch.bfloat16)
            lse = torch.empty(B, H, S, device=cuda)
            call("hct_attention_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, S, H, hd, stream_ptr(cuda))
            torch.cuda.synchronize()
            res[fold] = (out.float(), lse)
    finally:
        lib().hct_attention_set_tail_key(3)
    for fold in (1, 3):
        assert _rel(res[fold][0], res[0][0]) < 2e-3, fold
        assert _rel(res[fold][0][:, -1], res[0][0][:, -1]) < 6e-3, fold          # the last row on its own
        assert (res[fold][1] - res[0][1]).abs().max().item() < 1e-5, fold


def _attention_case(cuda, B, S, H, hd):
    from headct_foundation_b200._cabi import call, stream_ptr
    D = H * hd
    g = torch.Generator(device="cuda").manual_seed(S * 3 + H)
    qkv = torch.randn(B, S, 3 * D, device=cuda, generator=g).bfloat16()
    out = torch.empty(B, S, D, device=cuda, dtype=torch.bfloat16)
    lse = torch.empty(B, H, S, device=cuda)
    call("hct_attention_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, S, H, hd, stream_ptr(cuda))
    q, k, v = qkv.float().view(B, S, 3, H, hd).permute(2, 0, 3, 1, 4)
    q = q.clone().requires_grad_(True); k = k.clone().requires_grad_(True); v = v.clone().requires_grad_(True)
    att = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    ref = (att.softmax(-1) @ v).transpose(1, 2).reshape(B, S, D)
    assert _rel(out.float(), ref.detach()) < 8e-3
    assert _rel(lse, torch.logsumexp(att.detach(), -1)) < 1e-4
    do = torch.randn(B, S, D, device=cuda, generator=g).bfloat16()
    ref.backward(do.float())
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(B, H, S, device=cuda)
    # the entry point with the qkv-bias gradient: column sums of the dqkv it writes, added to a zeroed fp32 vector
    bias = torch.zeros(3 * D, device=cuda)
    call("hct_attention_bwd_bias", qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(),
         delta.data_ptr(), bias.data_ptr(), B, S, H, hd, stream_ptr(cuda))
    want = dqkv.double().sum((0, 1))
    assert (bias.double() - want).abs().max().item() < 1e-4 * dqkv.double().abs().sum((0, 1)).max().item() + 1e-6
    dref = torch.stack([q.grad, k.grad, v.grad]).permute(1, 3, 0, 2, 4).reshape(B, S, 3 * D)
    d = dqkv.float().view(B, S, 3, D)
    dr = dref.view(B, S, 3, D)
    for i, name in enumerate("qkv"):
        err = (d[:, :, i] - dr[:, :, i]).double().norm().item()
        assert err < 1.5e-2 * dr[:, :, i].double().norm().item() + 1e-5, name   # S = 1: dq and dk are exactly zero
    r0 = S // 128 * 128
    if 0 < r0 < S:          # the rows behind the last full tile on their own (a sliver of the norms above)
        for i, name in enumerate("qkv"):
            err = (d[:, r0:, i] - dr[:, r0:, i]).double().norm().item()
            assert err < 1.5e-2 * dr[:, r0:, i].double().norm().item() + 1e-5, ("tail", name)


# ------------------------------------------------------------------ masking (bit exact)
@pytest.mark.parametrize("N,L,ratio,ties", [(4, 512, 0.75, False), (7, 512, 0.75, True), (3, 64, 0.75, True),
                                            (2, 100, 0.6, False), (1, 8, 0.5, True), (2, 2000, 0.9, True)])
def test_mask_indices_bit_exact(cuda, HF, N, L, ratio, ties):
    from oracle import headct_oracle as O, synth
    noise = synth.noise(N, L, seed=L + N, ties=ties)
    ids_keep_o, ids_restore_o, mask_o, keep = O.masking_indices(noise, ratio)
    ids_restore, ids_keep, mask = HF.mask_indices(noise.to(cuda), keep)
    assert ids_restore.dtype == torch.int64 and ids_keep.dtype == torch.int64 and mask.dtype == torch.float32
    assert torch.equal(ids_restore.cpu(), ids_restore_o)
    assert torch.equal(ids_keep.cpu(), ids_keep_o)
    assert torch.equal(mask.cpu(), mask_o)
    # the library op the reference itself executes on a GPU (torch.argsort on CUDA), same tie rule
    ids_shuffle = torch.argsort(noise.to(cuda), dim=1)
    assert torch.equal(torch.argsort(ids_shuffle, dim=1), ids_restore)


def test_gather_scatter_tokens(cuda, HF):
    x = torch.randn(3, 64, 96, device=cuda, requires_grad=True)
    ids = torch.stack([torch.randperm(64, device=cuda)[:16] for _ in range(3)])
    y = HF.GatherTokensFn.apply(x, ids)
    ref = torch.gather(x.detach(), 1, ids[:, :, None].expand(-1, -1, 96))
    assert torch.equal(y.detach(), ref)
    g = torch.randn_like(y)
    y.backward(g)
    want = torch.zeros_like(x).scatter_(1, ids[:, :, None].expand(-1, -1, 96), g)
    assert torch.equal(x.grad, want)


# ------------------------------------------------------------------ patchify / window / loss / assemble
def test_patchify_matches_oracle(cuda, HF):
    from headct_foundation_b200._cabi import call, stream_ptr
    from oracle import headct_oracle as O, synth
    x = synth.volume(2, 3, 48, 3)
    cols_o = O.im2col_patches(x, (12, 12, 12))
    xc = x.to(cuda)
    cols = torch.empty(2 * 64, 5184, device=cuda, dtype=torch.bfloat16)
    idx = torch.empty(128, device=cuda, dtype=torch.int32)
    call("hct_patchify", xc.data_ptr(), cols.data_ptr(), None, idx.data_ptr(), 2, 3, 48, 48, 48, 12, 64, stream_ptr(cuda))
    assert torch.equal(cols.float().cpu().view(2, 64, 5184), cols_o.bfloat16().float())
    ids = torch.stack([torch.randperm(64)[:16] for _ in range(2)]).to(cuda)
    cols2 = torch.empty(2 * 16, 5184, device=cuda, dtype=torch.bfloat16)
    idx2 = torch.empty(32, device=cuda, dtype=torch.int32)
    call("hct_patchify", xc.data_ptr(), cols2.data_ptr(), ids.data_ptr(), idx2.data_ptr(), 2, 3, 48, 48, 48, 12, 16,
         stream_ptr(cuda))
    want = torch.gather(cols_o, 1, ids.cpu()[:, :, None].expand(-1, -1, 5184)).bfloat16().float()
    assert torch.equal(cols2.float().cpu().view(2, 16, 5184), want)
    assert torch.equal(idx2.cpu().long().view(2, 16), ids.cpu())


def test_window_scale_stack(cuda, HF):
    from oracle import headct_oracle as O, synth
    hu = synth.hu_volume(2, 16, 9)
    want = O.window_scale_stack(hu)
    got = HF.window_scale_stack(hu.to(cuda))
    assert got.shape == (2, 3, 16, 16, 16)
    assert torch.equal(got.cpu(), want)          # fp32 path is bit exact (same division, same clamp)
    got16 = HF.window_scale_stack(hu.to(cuda).to(torch.int16), out_dtype=torch.bfloat16)
    assert torch.equal(got16.float().cpu(), want.bfloat16().float())
    gold = np.load("tests/golden/misc.npz")
    g = HF.window_scale_stack(torch.from_numpy(gold["window_hu"]).to(cuda))
    assert torch.equal(g.cpu(), torch.from_numpy(gold["window_out"]))


@pytest.mark.parametrize("norm_pix", [False, True])
def test_mae_loss_fwd_bwd(cuda, HF, norm_pix):
    from oracle import headct_oracle as O, synth
    x = synth.volume(2, 3, 48, 4)
    L, P = 64, 5184
    g = torch.Generator().manual_seed(1)
    pred = (torch.randn(2, L + 1, P, generator=g) * 0.5).bfloat16()
    mask = (torch.rand(2, L, generator=g) < 0.75).float()
    pr = pred[:, 1:].float().clone().requires_grad_(True)
    want = O.mae_loss(x, pr, mask, (12, 12, 12), norm_pix)
    want.backward(torch.tensor(3.0))
    pc = pred.to(cuda).requires_grad_(True)
    loss = HF.MaeLossFn.apply(pc, x.to(cuda), mask.to(cuda), 12, norm_pix, False, 1)
    assert abs(loss.item() - want.item()) / want.item() < 1e-5
    (loss * 3.0).backward()
    assert torch.all(pc.grad[:, 0] == 0)
    assert _rel(pc.grad[:, 1:].float().cpu(), pr.grad) < 5e-3


def test_decoder_assemble_fwd_bwd(cuda, HF):
    from oracle import synth
    N, L, keep, D = 3, 64, 16, 96
    g = torch.Generator().manual_seed(2)
    y = torch.randn(N, keep + 1, D, generator=g).bfloat16()
    noise = synth.noise(N, L, 3)
    ids_restore = torch.argsort(torch.argsort(noise, 1), 1)
    mtok = torch.randn(1, 1, D, generator=g); cls = torch.randn(1, 1, D, generator=g); pos = torch.randn(1, L, D, generator=g)
    yr = y.float().requires_grad_(True); mr = mtok.clone().requires_grad_(True); cr = cls.clone().requires_grad_(True)
    pool = torch.cat([yr[:, 1:], mr.expand(N, L - keep, D)], 1)
    body = torch.gather(pool, 1, ids_restore[:, :, None].expand(-1, -1, D))
    ref = torch.cat([yr[:, :1], body], 1) + torch.cat([cr, pos], 1)
    yc = y.to(cuda).requires_grad_(True); mc = mtok.to(cuda).requires_grad_(True); cc = cls.to(cuda).requires_grad_(True)
    out = HF.DecoderAssembleFn.apply(yc, ids_restore.to(cuda), mc, cc, pos.to(cuda))
    assert _rel(out.cpu(), ref.detach()) < 1e-6
    go = torch.randn(N, L + 1, D, generator=g)
    ref.backward(go); out.backward(go.to(cuda))
    assert _rel(yc.grad.float().cpu(), yr.grad) < 5e-3
    assert _rel(mc.grad.cpu(), mr.grad) < 1e-5 and _rel(cc.grad.cpu(), cr.grad) < 1e-5


def test_colsum_broadcast_reduce(cuda, HF):
    from headct_foundation_b200._cabi import call, stream_ptr
    x = torch.randn(1000, 2304, device=cuda)
    assert _rel(HF.colsum(x, 2304), x.sum(0)) < 1e-5
    assert _rel(HF.colsum(x.bfloat16(), 2304), x.bfloat16().float().sum(0)) < 1e-5
    src = torch.randn(5, 96, device=cuda)
    dst = torch.zeros(4, 69, 96, device=cuda)
    call("hct_broadcast_rows", src.data_ptr(), dst.data_ptr(), 4, 5, 69, 0, 96, stream_ptr(cuda))
    assert torch.equal(dst[:, :5], src.expand(4, 5, 96)) and torch.all(dst[:, 5:] == 0)
    g = torch.randn(4, 69, 96, device=cuda)
    out = torch.zeros(5, 96, device=cuda)
    call("hct_reduce_rows", g.data_ptr(), out.data_ptr(), 4, 5, 69, 0, 96, stream_ptr(cuda))
    assert _rel(out, g[:, :5].sum(0)) < 1e-6


# ------------------------------------------------------------------ DINO kernels
def test_dino_loss_and_grad(cuda, HF):
    from oracle import headct_oracle as O
    B, ncrops, K = 3, 4, 4096
    g = torch.Generator().manual_seed(3)
    s = torch.randn(ncrops * B, K, generator=g); t = torch.randn(2 * B, K, generator=g); c = torch.randn(1, K, generator=g) * 0.1
    sr = s.clone().requires_grad_(True)
    want = O.dino_loss(sr, t, c, ncrops=ncrops, teacher_temp=0.04)
    want.backward(torch.tensor(2.0))
    sc = s.to(cuda).requires_grad_(True)
    loss = HF.DinoLossFn.apply(sc, t.to(cuda), c.to(cuda), ncrops, 0.1, 0.04)
    assert abs(loss.item() - want.item()) / abs(want.item()) < 1e-4
    (loss * 2.0).backward()
    assert _rel(sc.grad.cpu(), sr.grad) < 8e-3
    center = c.to(cuda).clone()
    HF.center_update(center, t.to(cuda), 0.9)
    assert _rel(center.cpu(), O.dino_center_update(c, t)) < 1e-5


def test_l2norm_weightnorm(cuda, HF):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(10, 256, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = torch.nn.functional.normalize(xr, dim=-1)
    xc = x.to(cuda).requires_grad_(True)
    y = HF.L2NormFn.apply(xc)
    assert _rel(y.float().cpu(), yr.detach()) < 4e-3
    go = torch.randn(10, 256, generator=g)
    yr.backward(go); y.backward(go.to(cuda).bfloat16())
    assert _rel(xc.grad.cpu(), xr.grad) < 1.5e-2
    # weight_norm with a TRAINABLE gain (DINOHead(norm_last_layer=False), dino_head.py:28-29): dv and dg
    v = torch.randn(512, 64, generator=g) * 0.02; gg = 0.5 + torch.rand(512, 1, generator=g)
    vr = v.clone().requires_grad_(True); gr = gg.clone().requires_grad_(True)
    w = vr * (gr / vr.norm(dim=1, keepdim=True))
    xin = torch.randn(6, 64, generator=g).bfloat16()
    logits_r = xin.float() @ w.t()
    vc = v.to(cuda).requires_grad_(True); gc = gg.to(cuda).requires_grad_(True)
    logits = HF.WeightNormLinearFn.apply(xin.to(cuda), gc, vc)
    assert _rel(logits.cpu(), logits_r.detach()) < 5e-3
    gl = torch.randn(6, 512, generator=g)
    logits_r.backward(gl); logits.backward(gl.to(cuda))
    assert _rel(vc.grad.cpu(), vr.grad) < 1.5e-2
    assert gc.grad is not None and _rel(gc.grad.cpu(), gr.grad) < 1.5e-2
    # frozen gain (norm_last_layer=True): no dg is produced
    gc2 = gg.to(cuda); vc2 = v.to(cuda).requires_grad_(True)
    HF.WeightNormLinearFn.apply(xin.to(cuda), gc2, vc2).backward(gl.to(cuda))
    assert _rel(vc2.grad.cpu(), vr.grad) < 1.5e-2 and gc2.grad is None


def test_ema_and_adamw_multi(cuda, HF):
    from headct_foundation_b200._cabi import call, stream_ptr
    from oracle import headct_oracle as O
    g = torch.Generator().manual_seed(6)
    shapes = [(768,), (3, 5, 7), (2304, 768), (1,), (130, 3)]
    sp = [torch.randn(*s, generator=g) for s in shapes]; tp = [torch.randn(*s, generator=g) for s in shapes]
    tr = [t.clone() for t in tp]
    O.ema_update(tr, sp, 0.996)
    tc = [torch.nn.Parameter(t.to(cuda)) for t in tp]
    w = HF.w16(tc[2])                                     # a cached bf16 copy exists for the 2-D weight
    v0 = tc[2]._version
    HF.ema_update([s.to(cuda) for s in sp], tc, 0.996)
    for a, b in zip(tc, tr):
        assert _rel(a.detach().cpu(), b) < 1e-6
    assert tc[2]._version > v0                            # raw-pointer update is visible to version-keyed caches
    assert HF.w16(tc[2]) is w and torch.equal(w, tc[2].detach().bfloat16())     # ... and the copy was refreshed in place
    # per-parameter clip + AdamW, two steps
    ps = [torch.randn(*s, generator=g) for s in shapes]
    gs = [torch.randn(*s, generator=g) * (10.0 if i % 2 else 0.01) for i, s in enumerate(shapes)]
    pr = [p.clone() for p in ps]; gr = [x.clone() for x in gs]
    mr = [torch.zeros_like(p) for p in ps]; vr = [torch.zeros_like(p) for p in ps]
    pc = [p.to(cuda) for p in ps]; gc = [x.to(cuda) for x in gs]
    mc = [torch.zeros_like(p) for p in pc]; vc = [torch.zeros_like(p) for p in pc]
    # bf16 shadows (the GEMM operand copies) for every second tensor: the AdamW launch must keep them current
    sh = [torch.zeros(p.shape, dtype=torch.bfloat16, device=cuda) if i % 2 == 0 else None for i, p in enumerate(pc)]
    table = torch.tensor([[p.data_ptr(), x.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), 0 if h is None else h.data_ptr(), 0]
                          for p, x, m, v, h in zip(pc, gc, mc, vc, sh)], dtype=torch.int64, device=cuda)
    norms = torch.empty(len(pc), device=cuda)
    for step in (1, 2):
        gclip = [x.clone() for x in gr]
        O.clip_per_param(gclip, 3.0)
        for p, x, m, v in zip(pr, gclip, mr, vr):
            O.adamw_step(p, x, m, v, step, lr=1e-2, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.05)
        call("hct_grad_norms_multi", table.data_ptr(), len(pc), norms.data_ptr(), stream_ptr(cuda))
        call("hct_adamw_multi", table.data_ptr(), len(pc), norms.data_ptr(), 3.0, 1e-2, 0.9, 0.95, 1e-8, 0.05, step,
             stream_ptr(cuda))
    for a, b in zip(pc, pr):
        assert _rel(a.cpu(), b) < 1e-5
    for a, h in zip(pc, sh):
        if h is not None:
            assert torch.equal(h, a.bfloat16())


def test_fused_adamw_per_parameter_step_matches_torch(cuda, HF):
    """torch.optim.AdamW bias-corrects every parameter with its own step count.  DINO's cancel_gradients_last_layer
    (misc.py:366-371, engine_pretrain_dino.py:95) sets last_layer grads to None during the first epoch, so that tensor
    joins later with zero moments and step 0: FusedAdamW must give it ITS bias correction (round 1 used the group's),
    and a tensor that drops out for a few steps must resume with its own count.  Compared with stock AdamW every step."""
    from headct_foundation_b200.optim import FusedAdamW
    g = torch.Generator().manual_seed(11)
    shapes = [(64, 32), (32,), (128, 16), (7,)]
    init = [torch.randn(*s, generator=g) for s in shapes]
    ours = [torch.nn.Parameter(t.clone().to(cuda)) for t in init]
    ref = [torch.nn.Parameter(t.clone().to(cuda)) for t in init]
    kw = dict(lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.04)
    fo = FusedAdamW(ours, clip_grad=0.0, **kw)
    ro = torch.optim.AdamW(ref, **kw)
    frozen_until = {2: 6}                 # tensor 2 ("last_layer") has no gradient for the first 6 steps
    pause = {1: range(9, 12)}             # tensor 1 drops out for steps 9..11 and comes back
    for step in range(16):
        for i, (po, pr) in enumerate(zip(ours, ref)):
            gr = torch.randn(*shapes[i], generator=g).to(cuda)
            off = step < frozen_until.get(i, 0) or step in pause.get(i, ())
            po.grad = None if off else gr.clone()
            pr.grad = None if off else gr.clone()
        fo.step(); ro.step()
        for i, (po, pr) in enumerate(zip(ours, ref)):
            assert _rel(po.detach().cpu(), pr.detach().cpu()) < 2e-6, (step, i)
    assert int(fo.state[ours[2]]["step"]) == 10 and int(fo.state[ours[1]]["step"]) == 13 and int(fo.state[ours[0]]["step"]) == 16
    # the first tensor of the table may itself be the late one (table keyed on "params with a grad")
    ours2 = [torch.nn.Parameter(t.clone().to(cuda)) for t in init[:2]]
    ref2 = [torch.nn.Parameter(t.clone().to(cuda)) for t in init[:2]]
    fo2, ro2 = FusedAdamW(ours2, **kw), torch.optim.AdamW(ref2, **kw)
    for step in range(8):
        for i, (po, pr) in enumerate(zip(ours2, ref2)):
            gr = torch.randn(*shapes[i], generator=g).to(cuda)
            off = i == 0 and step < 4
            po.grad = None if off else gr.clone()
            pr.grad = None if off else gr.clone()
        fo2.step(); ro2.step()
        for po, pr in zip(ours2, ref2):
            assert _rel(po.detach().cpu(), pr.detach().cpu()) < 2e-6, step
