"""Same-box comparator: stock torch (cuBLAS / SDPA / eager ops) running the reference algorithm under bf16 autocast on the
B200 against this repo's kernels -- the "kernel set to beat" of SURVEY.md section 0 / BASELINE.md section 4.

    python tools/stock_torch_compare.py [--out profiles/r02_vs_stock_torch.json] [--batch 64]

Three comparisons, all timed with CUDA events after warm-up:
  step   full MAE pre-training step (fwd + loss + bwd) of the mae_HeadCT.yaml model: the oracle restatement (plain torch ops
         on a flat state_dict, F.scaled_dot_product_attention for attentionblock.py:61) under torch.autocast(bf16), vs
         headct_foundation_b200.MaskedAutoencoderViT, same batch, same weights.  The oracle is the reference's algorithm
         op for op (pinned to it by oracle/gen_golden.py); the reference package itself cannot travel to the GPU box.
  sdpa   F.scaled_dot_product_attention (bf16) fwd / fwd+bwd vs hct_attention_fwd / _bwd at the encoder, decoder and
         DINO shapes.
  gemm   torch.matmul (cuBLAS bf16) vs hct_gemm_bf16 on the GEMM shapes of one B = 256 step (plain epilogue on both
         sides, so the fused epilogues are not credited here).
This tool is a measurement aid: the oracle is used only as the thing compared against."""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

from oracle import headct_oracle as O, synth

dev = torch.device("cuda")


def timeit(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def sdpa_block(x, sd, pre, heads):
    """attentionblock.py:51-66 with the fused SDPA the reference calls (the oracle materialises the score matrix)."""
    B, S, C = x.shape
    qkv = F.linear(x, sd[pre + "qkv.weight"], sd.get(pre + "qkv.bias"))
    qkv = qkv.reshape(B, S, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    y = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2]).transpose(1, 2).reshape(B, S, C)
    return F.linear(y, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def stock_block(x, sd, pre, heads):
    h = F.layer_norm(x, (x.shape[-1],), sd[pre + "att_norm.weight"], sd[pre + "att_norm.bias"], 1e-5)
    x = x + sdpa_block(h, sd, pre + "attn.", heads)
    h = F.layer_norm(x, (x.shape[-1],), sd[pre + "ffn_norm.weight"], sd[pre + "ffn_norm.bias"], 1e-5)
    h = F.linear(F.gelu(F.linear(h, sd[pre + "mlp.linear1.weight"], sd[pre + "mlp.linear1.bias"])),
                 sd[pre + "mlp.linear2.weight"], sd[pre + "mlp.linear2.bias"])
    return x + h


def stock_mae_step(sd, x, noise, cfg, autocast=True):
    """mae.py:220-317 with stock torch modules' functional forms (Conv3d, LayerNorm, Linear, SDPA, GELU) under autocast."""
    p = cfg["patch_size"]
    with torch.autocast(x.device.type, dtype=torch.bfloat16, enabled=autocast):
        t = F.conv3d(x, sd["patch_embedding.patch_embeddings.weight"], sd["patch_embedding.patch_embeddings.bias"], stride=p)
        t = t.flatten(2).transpose(-1, -2) + sd["patch_embedding.position_embeddings"]
        t, mask, ids_restore, _ = O.random_masking(t, noise, cfg["mask_ratio"])
        t = torch.cat([sd["cls_token"].expand(t.shape[0], -1, -1).to(t.dtype), t], dim=1)
        for i in range(cfg["encoder_depth"]):
            t = stock_block(t, sd, f"blocks.{i}.", cfg["encoder_num_heads"])
        t = F.layer_norm(t, (t.shape[-1],), sd["norm.weight"], sd["norm.bias"], 1e-5)
        pred = _stock_decoder(sd, t, ids_restore, cfg)
        loss = O.mae_loss(x, pred.float(), mask, (p, p, p), cfg["norm_pix_loss"])
    return loss


def _stock_decoder(sd, latent, ids_restore, cfg):
    B, L = ids_restore.shape
    y = F.linear(latent, sd["decoder_embed.weight"], sd.get("decoder_embed.bias"))
    C = y.shape[2]
    pool = torch.cat([y[:, 1:], sd["mask_token"].expand(B, L + 1 - y.shape[1], C).to(y.dtype)], dim=1)
    body = torch.gather(pool, 1, ids_restore[:, :, None].expand(-1, -1, C))
    y = torch.cat([y[:, :1], body], dim=1) + torch.cat([sd["decoder_cls_token"], sd["decoder_pos_embed"]], dim=1)
    for i in range(cfg["decoder_depth"]):
        y = stock_block(y, sd, f"decoder_blocks.{i}.", cfg["decoder_num_heads"])
    y = F.layer_norm(y, (C,), sd["decoder_norm.weight"], sd["decoder_norm.bias"], 1e-5)
    return F.linear(y, sd["decoder_pred.weight"], sd.get("decoder_pred.bias"))[:, 1:]


def compare_step(batch):
    cfg = synth.MAE_FULL
    sd = synth.to_device(synth.mae_state_dict(cfg, seed=4), dev)
    sdg = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    torch.manual_seed(0)
    out = {}
    B = batch
    while B >= 8:
        try:
            x = torch.rand(B, 3, 96, 96, 96, device=dev)
            noise = torch.rand(B, 512, device=dev)

            def stock():
                for v in sdg.values():
                    v.grad = None
                stock_mae_step(sdg, x, noise, cfg).backward()
            ms_stock = timeit(stock, warmup=2, iters=4)
            loss_stock = float(stock_mae_step(sdg, x, noise, cfg))
            break
        except torch.OutOfMemoryError:
            torch.cuda.empty_cache()
            B //= 2
    model = H.MaskedAutoencoderViT(**cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).train()
    model.noise_override = noise

    def ours():
        model.zero_grad(set_to_none=True)
        model(x)[0].backward()
    ms_ours = timeit(ours, warmup=2, iters=4)
    loss_ours = float(model(x)[0])
    gf = 282.133 * B
    out = {"batch": B, "stock_torch_ms": ms_stock, "ours_ms": ms_ours, "speedup": ms_stock / ms_ours,
           "stock_volumes_per_s": B / ms_stock * 1e3, "ours_volumes_per_s": B / ms_ours * 1e3,
           "stock_tflops": gf / ms_stock, "ours_tflops": gf / ms_ours, "loss_stock_bf16_autocast": loss_stock,
           "loss_ours": loss_ours,
           "what": "MAE ViT-B 3D fwd+loss+bwd (no optimizer), mae_HeadCT.yaml shape; stock = torch "
                   f"{torch.__version__} functional ops (conv3d, layer_norm, linear, SDPA, gelu) under autocast(bf16)"}
    del model, sdg
    torch.cuda.empty_cache()
    return out


def compare_sdpa():
    rows = []
    for name, (B, S, Hh, hd) in {"decoder": (256, 513, 16, 48), "encoder": (256, 129, 12, 64), "dino": (64, 517, 12, 64)}.items():
        D = Hh * hd
        qkv = torch.randn(B, S, 3 * D, device=dev).bfloat16()
        q5 = qkv.view(B, S, 3, Hh, hd).permute(2, 0, 3, 1, 4)
        q, k, v = (t.detach().requires_grad_(True) for t in (q5[0], q5[1], q5[2]))      # strided views, as the reference passes them
        do = torch.randn(B, S, D, device=dev).bfloat16()
        f_stock = timeit(lambda: F.scaled_dot_product_attention(q, k, v))

        def fb():
            q.grad = k.grad = v.grad = None
            o = F.scaled_dot_product_attention(q, k, v)
            o.backward(do.view(B, S, Hh, hd).transpose(1, 2))
        fb_stock = timeit(fb)
        out = torch.empty(B, S, D, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(B, Hh, S, device=dev)
        dqkv = torch.empty_like(qkv)
        delta = torch.empty(B, Hh, S, device=dev)
        st = stream_ptr(dev)
        f_ours = timeit(lambda: call("hct_attention_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, S, Hh, hd, st))
        b_ours = timeit(lambda: call("hct_attention_bwd", qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(),
                                     dqkv.data_ptr(), delta.data_ptr(), B, S, Hh, hd, st))
        fl = 4.0 * B * Hh * S * S * hd
        rows.append({"shape": name, "B": B, "S": S, "heads": Hh, "head_dim": hd, "sdpa_fwd_ms": f_stock,
                     "sdpa_bwd_ms": fb_stock - f_stock, "ours_fwd_ms": f_ours, "ours_bwd_ms": b_ours,
                     "fwd_speedup": f_stock / f_ours, "bwd_speedup": (fb_stock - f_stock) / b_ours,
                     "ours_fwd_tflops": fl / f_ours / 1e9, "sdpa_fwd_tflops": fl / f_stock / 1e9,
                     "ours_bwd_tflops": 2 * fl / b_ours / 1e9, "sdpa_bwd_tflops": 2 * fl / (fb_stock - f_stock) / 1e9})
    return rows


def compare_gemm():
    rows = []
    shapes = []
    for tag, M in (("dec", 131328), ("enc", 33024)):
        shapes += [(f"{tag} qkv fwd", M, 2304, 768, "nt"), (f"{tag} proj fwd", M, 768, 768, "nt"),
                   (f"{tag} fc1 fwd", M, 3072, 768, "nt"), (f"{tag} fc2 fwd", M, 768, 3072, "nt"),
                   (f"{tag} fc2 dgrad", M, 3072, 768, "nn"), (f"{tag} fc1 dgrad", M, 768, 3072, "nn"),
                   (f"{tag} qkv dgrad", M, 768, 2304, "nn"), (f"{tag} fc1 wgrad", 3072, 768, M, "tn"),
                   (f"{tag} fc2 wgrad", 768, 3072, M, "tn"), (f"{tag} qkv wgrad", 2304, 768, M, "tn")]
    shapes += [("pred fwd", 131328, 5184, 768, "nt"), ("pred dgrad", 131328, 768, 5184, "nn"),
               ("pred wgrad", 5184, 768, 131328, "tn"), ("embed fwd", 32768, 768, 5184, "nt")]
    for name, M, N, K, kind in shapes:
        if kind == "nt":      # C = A[M,K] B[N,K]^T
            A = torch.randn(M, K, device=dev).bfloat16(); Bm = torch.randn(N, K, device=dev).bfloat16()
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            stock = lambda: torch.matmul(A, Bm.t(), out=out)
            ours = lambda: HF.gemm(A, Bm, M=M, N=N, K=K, lda=K, ldb=K, out=out, ldo=N, epi=HF.EPI_BF16)
        elif kind == "nn":    # C = A[M,K] B[K,N]
            A = torch.randn(M, K, device=dev).bfloat16(); Bm = torch.randn(K, N, device=dev).bfloat16()
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            stock = lambda: torch.matmul(A, Bm, out=out)
            ours = lambda: HF.gemm(A, Bm, M=M, N=N, K=K, lda=K, ldb=N, b_mn=True, out=out, ldo=N, epi=HF.EPI_BF16)
        else:                 # C[M,N] = A[K,M]^T B[K,N], fp32 out (weight gradient)
            A = torch.randn(K, M, device=dev).bfloat16(); Bm = torch.randn(K, N, device=dev).bfloat16()
            out32 = torch.zeros(M, N, device=dev)
            out16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            stock = lambda: torch.matmul(A.t(), Bm, out=out16)          # cuBLAS bf16 out (autocast wgrads are bf16 -> fp32 copy)
            ours = lambda: HF.gemm(A, Bm, M=M, N=N, K=K, lda=M, ldb=N, a_mn=True, b_mn=True, out=out32, ldo=N,
                                   epi=HF.EPI_ATOMIC_F32)
        ms_s, ms_o = timeit(stock, iters=6), timeit(ours, iters=6)
        fl = 2.0 * M * N * K
        rows.append({"shape": name, "M": M, "N": N, "K": K, "cublas_ms": ms_s, "ours_ms": ms_o, "cublas_tflops": fl / ms_s / 1e9,
                     "ours_tflops": fl / ms_o / 1e9, "speedup": ms_s / ms_o})
        del A, Bm
    return rows


if __name__ == "__main__":
    import headct_foundation_b200 as H
    from headct_foundation_b200 import functional as HF
    from headct_foundation_b200._cabi import call, stream_ptr
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--only", default="step,sdpa,gemm")
    a = ap.parse_args()
    res = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
    only = a.only.split(",")
    if "sdpa" in only:
        res["sdpa"] = compare_sdpa()
    if "gemm" in only:
        res["gemm"] = compare_gemm()
        tot_s = sum(r["cublas_ms"] for r in res["gemm"]); tot_o = sum(r["ours_ms"] for r in res["gemm"])
        res["gemm_total"] = {"cublas_ms": tot_s, "ours_ms": tot_o, "speedup": tot_s / tot_o}
    if "step" in only:
        res["step"] = compare_step(a.batch)
    txt = json.dumps(res, indent=1)
    print(txt)
    if a.out:
        with open(os.path.join(ROOT, a.out), "w") as f:
            f.write(txt + "\n")
