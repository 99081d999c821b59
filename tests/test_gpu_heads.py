"""GPU parity for the downstream rows: RMSNorm blocks, LoRA q/v adapters (with the reference's
reshape quirk), the attentive-pooling classifier and its BatchNorm-over-tokens, against golden vectors produced by the
unmodified reference (oracle/gen_golden.py) and fp32 torch restatements.  Tolerances as in test_gpu_models.py:
features cosine >= 0.999, gradients cosine > 0.99 and norms within 5 %."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def HF():
    from headct_foundation_b200 import functional
    return functional


# ------------------------------------------------------------------ kernels
@pytest.mark.parametrize("rows,dim", [(257, 768), (1000, 192), (33, 1024)])
def test_rmsnorm_fwd_bwd(cuda, HF, rows, dim):
    g = torch.Generator(device="cuda").manual_seed(rows + dim)
    x = torch.randn(rows, dim, device=cuda, generator=g) * 2 + 0.5
    w = torch.randn(dim, device=cuda, generator=g) * 0.1 + 1
    dy = torch.randn(rows, dim, device=cuda, generator=g)
    dres = torch.randn(rows, dim, device=cuda, generator=g)
    eps = 1e-6
    y, mean, rstd = HF.layernorm_fwd(x, w, None, eps, False, True)
    assert mean is None
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    ref = xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + eps) * wr            # layers.py:40,52-53
    assert _rel(y, ref.detach()) < 1e-5
    y16, _, _ = HF.layernorm_fwd(x, w, None, eps, True, False)
    assert _rel(y16.float(), ref.detach()) < 5e-3
    ref.backward(dy)
    for dy_in in (dy, dy.bfloat16()):
        dx, dx16, dg, db = HF.layernorm_bwd(dy_in, x, w, None, rstd, dres, True)
        assert db is None
        tol = 1e-4 if dy_in.dtype == torch.float32 else 6e-3
        assert _rel(dx, xr.grad + dres) < tol and _rel(dg, wr.grad) < tol
        assert _rel(dx16.float(), xr.grad + dres) < 6e-3


@pytest.mark.parametrize("B,S,H,hd", [(2, 69, 3, 64), (3, 513, 16, 48), (1, 17, 4, 32)])
def test_lora_shuffle_and_adjoint(cuda, HF, B, S, H, hd):
    from headct_foundation_b200._cabi import call, stream_ptr
    D = H * hd
    g = torch.Generator(device="cuda").manual_seed(S)
    qkv = torch.randn(B, S, 3 * D, device=cuda, generator=g).bfloat16()
    lq = torch.randn(B, S, D, device=cuda, generator=g).bfloat16()
    lv = torch.randn(B, S, D, device=cuda, generator=g).bfloat16()
    out = HF.LoraAddFn.apply(qkv, lq, lv, H)
    q5 = qkv.float().view(B, S, 3, H, hd).clone()
    q5[:, :, 0] += lq.float().reshape(B, H, S, hd).permute(0, 2, 1, 3)             # attentionblock.py:58: reshape, no transpose
    q5[:, :, 2] += lv.float().reshape(B, H, S, hd).permute(0, 2, 1, 3)
    assert torch.equal(out.view(B, S, 3, H, hd), q5.bfloat16())                    # one bf16 rounding of an exact fp32 sum
    dq = torch.randn(B, S, 3 * D, device=cuda, generator=g).bfloat16()
    dlq, dlv = torch.empty_like(lq), torch.empty_like(lv)
    call("hct_lora_shuffle", dq.data_ptr(), dlq.data_ptr(), dlv.data_ptr(), B, S, H, hd, 1, stream_ptr(cuda))
    d5 = dq.view(B, S, 3, H, hd)
    assert torch.equal(dlq, d5[:, :, 0].permute(0, 2, 1, 3).reshape(B, S, D))       # pure data movement: bit exact
    assert torch.equal(dlv, d5[:, :, 2].permute(0, 2, 1, 3).reshape(B, S, D))


@pytest.mark.parametrize("rows,dim,training", [(4 * 513, 768, True), (1001, 192, True), (777, 96, False)])
def test_colnorm_fwd_bwd(cuda, HF, rows, dim, training):
    g = torch.Generator(device="cuda").manual_seed(rows)
    x = torch.randn(rows, dim, device=cuda, generator=g) * 1.7 + 0.4
    dy = torch.randn(rows, dim, device=cuda, generator=g)
    bn = torch.nn.BatchNorm1d(dim, affine=False, eps=1e-6).to(cuda).train(training)
    bn.running_mean.uniform_(-0.5, 0.5); bn.running_var.uniform_(0.5, 2.0)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    xr = x.clone().requires_grad_(True)
    ref = bn(xr)
    ref.backward(dy)
    xo = x.clone().requires_grad_(True)
    y = HF.ColNormFn.apply(xo, rm, rv, training, 1e-6, 0.1, False)
    assert _rel(y, ref.detach()) < 2e-5
    assert _rel(rm, bn.running_mean) < 1e-5 and _rel(rv, bn.running_var) < 1e-5
    y.backward(dy)
    assert _rel(xo.grad, xr.grad) < 2e-4
    y16 = HF.ColNormFn.apply(xo, rm.clone(), rv.clone(), training, 1e-6, 0.1, True)
    assert y16.dtype == torch.bfloat16 and _rel(y16.float(), ref.detach()) < 5e-3


@pytest.mark.parametrize("B,N,H,hd,nq", [(4, 513, 12, 64, 1), (3, 69, 3, 64, 1), (2, 33, 2, 48, 3), (2, 517, 4, 32, 8)])
def test_pool_attention_fwd_bwd(cuda, HF, B, N, H, hd, nq):
    C = H * hd
    g = torch.Generator(device="cuda").manual_seed(N + nq)
    cls = (torch.randn(nq, C, device=cuda, generator=g) * 2).requires_grad_(True)
    kv = torch.randn(B, N, 2 * C, device=cuda, generator=g).bfloat16().requires_grad_(True)
    dout = torch.randn(B, nq, C, device=cuda, generator=g)
    scale = hd ** -0.5
    out = HF.PoolAttentionFn.apply(cls, kv, H, scale * scale)
    out.backward(dout)
    clsr = cls.detach().clone().requires_grad_(True)
    kvr = kv.detach().float().requires_grad_(True)
    q = clsr.view(1, nq, H, hd).expand(B, -1, -1, -1).permute(0, 2, 1, 3) * scale                  # classifier.py:86-87
    k, v = kvr.view(B, N, 2, H, hd).permute(2, 0, 3, 1, 4)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).reshape(B, nq, C)                # :93-95 (reshape quirk)
    # our op returns [B, nq, (H, hd)] = the transpose(1,2) layout; the reference's reshape of [B,H,nq,hd] equals it for
    # nq == 1 only, so compare through the explicit head-major view
    ref_t = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, nq, C)
    ref_t.backward(dout)
    assert _rel(out, ref_t.detach()) < 1e-4
    assert _rel(cls.grad, clsr.grad) < 2e-3
    assert _rel(kv.grad.float(), kvr.grad) < 6e-3
    if nq == 1:
        assert torch.allclose(ref, ref_t)


# ------------------------------------------------------------------ models
@pytest.mark.parametrize("name", ["vit_small_lora", "vit_small_rms"])
def test_vit_variants_features(cuda, name):
    import headct_foundation_b200 as H
    from oracle import synth
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = json.loads(str(gold["cfg"]))
    sd = synth.vit_state_dict(cfg, seed=int(gold["w_seed"]))
    m = H.ViT(**synth.resolve_norm(cfg, H.RMSNorm))
    m.load_state_dict(sd, strict=True)
    m = m.to(cuda).eval()
    x = synth.volume(int(gold["batch"]), cfg["in_chans"], cfg["img_size"], int(gold["x_seed"])).to(cuda)
    with torch.no_grad():
        y, hidden = m(x)
    nreg = cfg.get("num_register_tokens", 0)
    for b in range(y.shape[0]):
        assert _cos(y[b, 0].cpu(), torch.from_numpy(gold["cls"][b])) >= 0.999
        assert _cos(y[b, 1 + nreg:].mean(0).cpu(), torch.from_numpy(gold["pooled"][b])) >= 0.999
    assert _cos(y.cpu(), torch.from_numpy(gold["tokens"])) > 0.9995
    assert _rel(hidden[-1].norm(dim=-1).cpu(), torch.from_numpy(gold["hidden_norms"][-1])) < 1e-2


def test_lora_finetune_gradients(cuda):
    """TRAIN.LORA: only lora / bias / embeddings / norm parameters train (misc.py:349-359); their gradients match the
    reference's, frozen weights receive none (and their wgrad GEMMs are skipped)."""
    import headct_foundation_b200 as H
    from headct_foundation_b200 import _cabi
    from oracle import synth
    gold = np.load(os.path.join(GOLD, "vit_small_lora_grads.npz"))
    cfg = synth.VIT_SMALL_LORA
    m = H.ViT(**cfg)
    m.load_state_dict(synth.vit_state_dict(cfg, seed=6), strict=True)
    H.set_requires_grad_false(m, lora=True)
    assert sorted(n for n, p in m.named_parameters() if p.requires_grad) == json.loads(str(gold["trainable"]))
    m = m.to(cuda).train()
    x = synth.volume(2, cfg["in_chans"], cfg["img_size"], 5).to(cuda)
    y, _ = m(x)
    w = torch.from_numpy(np.random.default_rng(int(gold["out_weight_seed"])).standard_normal(tuple(y.shape)).astype(np.float32))
    n0 = _cabi.launch_count()
    (y * w.to(cuda)).sum().backward()
    lora_launches = _cabi.launch_count() - n0
    params = dict(m.named_parameters())
    for k in gold.files:
        if k.startswith("grad::"):
            gr, ref = params[k[6:]].grad.cpu(), torch.from_numpy(gold[k])
            assert _cos(gr, ref) > 0.99, (k, _cos(gr, ref))
            assert abs(gr.norm().item() - ref.norm().item()) < 0.05 * ref.norm().item() + 1e-6, k
    assert all(p.grad is None for p in params.values() if not p.requires_grad)
    # the same backward with every parameter trainable launches more kernels (the frozen weights' wgrad GEMMs)
    for p in m.parameters():
        p.requires_grad = True
    y, _ = m(x)
    n0 = _cabi.launch_count()
    (y * w.to(cuda)).sum().backward()
    assert _cabi.launch_count() - n0 > lora_launches


def test_attention_classifier(cuda):
    import headct_foundation_b200 as H
    from oracle import synth
    gold = np.load(os.path.join(GOLD, "attention_classifier.npz"))
    for tag in ("q1", "q3"):
        c = json.loads(str(gold[tag + "_cfg"]))
        clf = H.AttentionClassifier(c["dim"], 2, num_heads=c["heads"], qkv_bias=c["bias"], num_queries=c["nq"])
        clf.load_state_dict(synth.attention_classifier_state_dict(c["dim"], 2, num_queries=c["nq"], qkv_bias=c["bias"], seed=51))
        clf = clf.to(cuda).train()
        x = torch.from_numpy(gold[tag + "_x"]).to(cuda).requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.float16):          # the engine calls it under autocast (engine_downstream.py:78)
            logits = clf(x)
        (logits.float() * torch.from_numpy(gold[tag + "_wl"]).to(cuda)).sum().backward()
        ref = torch.from_numpy(gold[tag + "_logits_train"])
        assert _rel(logits.float().cpu(), ref) < 1e-2, (tag, logits, ref)
        assert _cos(x.grad.cpu(), torch.from_numpy(gold[tag + "_dx"])) > 0.99
        assert _cos(clf.cls_token.grad.cpu(), torch.from_numpy(gold[tag + "_dcls"])) > 0.99
        assert _cos(clf.wkv.weight.grad.cpu(), torch.from_numpy(gold[tag + "_dwkv"])) > 0.99
        assert abs(x.grad.norm().item() - np.linalg.norm(gold[tag + "_dx"])) < 0.05 * np.linalg.norm(gold[tag + "_dx"])
        assert _rel(clf.bn1.running_mean.cpu(), torch.from_numpy(gold[tag + "_bn1_mean"])) < 1e-4
        assert _rel(clf.bn1.running_var.cpu(), torch.from_numpy(gold[tag + "_bn1_var"])) < 1e-4
        assert int(clf.bn1.num_batches_tracked) == 1
        clf.eval()
        with torch.no_grad():
            le = clf(x.detach())
        assert _rel(le.float().cpu(), torch.from_numpy(gold[tag + "_logits_eval"])) < 1e-2


# ------------------------------------------------------------------ full shipped shapes
def test_lora_zero_init_equals_base_model_full_size(cuda):
    """vit_HeadCT_cq500.yaml shape with TRAIN.LORA: a freshly constructed adapter has lora_matrix_B == 0
    (attentionblock.py:18), so the LoRA model must reproduce the base model BIT FOR BIT -- the low-rank GEMMs and the
    reshape-add contribute exact zeros -- and its backward must leave the frozen weights without gradients."""
    import headct_foundation_b200 as H
    from headct_foundation_b200.configs import VIT_DOWNSTREAM
    torch.manual_seed(3)
    base = H.ViT(**VIT_DOWNSTREAM).to(cuda).eval()
    lora = H.ViT(**dict(VIT_DOWNSTREAM, lora=True)).to(cuda).eval()
    missing, unexpected = lora.load_state_dict(base.state_dict(), strict=False)
    assert not unexpected and all("lora" in k for k in missing)
    x = torch.rand(4, 3, 96, 96, 96, device=cuda)
    with torch.no_grad():
        yb, hb = base(x)
        yl, hl = lora(x)
    assert torch.equal(yb, yl) and all(torch.equal(a, b) for a, b in zip(hb, hl))
    H.set_requires_grad_false(lora, lora=True)
    lora.train()
    y, _ = lora(x)
    y[:, 0].square().mean().backward()
    named = dict(lora.named_parameters())
    assert all(p.grad is None for p in named.values() if not p.requires_grad)
    gB = named["blocks.11.attn.lora_v.lora_matrix_B"].grad
    assert gB is not None and torch.isfinite(gB).all() and gB.abs().max() > 0
    gA = named["blocks.11.attn.lora_v.lora_matrix_A"].grad          # dA = B^T (...) = 0 while B == 0
    assert gA is not None and gA.abs().max() == 0


def test_attention_classifier_full_shape_vs_oracle(cuda):
    """Attentive probe at the downstream shape (batch 64 x 513 tokens x 768, 12 heads): logits against the CPU oracle
    (1e-2 relative, bf16 kv path), train and eval mode; gradients reach tokens, queries and wkv."""
    import headct_foundation_b200 as H
    from oracle import headct_oracle as O, synth
    sd = synth.attention_classifier_state_dict(768, 2, num_queries=1, qkv_bias=False, seed=71)
    clf = H.AttentionClassifier(768, 2, num_heads=12, num_queries=1)
    clf.load_state_dict(sd)
    clf = clf.to(cuda).train()
    g = torch.Generator().manual_seed(72)
    x = torch.randn(64, 513, 768, generator=g) * 1.3 + 0.2
    xc = x.to(cuda).requires_grad_(True)
    logits = clf(xc)
    ref = O.attention_classifier(sd, x, 12, training=True)
    assert _rel(logits.float().cpu(), ref) < 1e-2
    w = torch.randn(64, 2, generator=g)
    (logits.float() * w.to(cuda)).sum().backward()
    xr = x.clone().requires_grad_(True)
    sdr = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    (O.attention_classifier(sdr, xr, 12, training=True) * w).sum().backward()
    assert _cos(xc.grad.cpu(), xr.grad) > 0.99
    assert _cos(clf.cls_token.grad.cpu(), sdr["cls_token"].grad) > 0.99
    assert _cos(clf.wkv.weight.grad.cpu(), sdr["wkv.weight"].grad) > 0.99
    clf.eval()
    sd_eval = {k: v.detach().cpu() for k, v in clf.state_dict().items()}
    with torch.no_grad():
        le = clf(xc.detach())
    assert _rel(le.float().cpu(), O.attention_classifier(sd_eval, x, 12, training=False)) < 1e-2
