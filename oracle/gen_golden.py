"""Generate tests/golden/* by EXECUTING THE UNMODIFIED REFERENCE (this container only).

    python -m oracle.gen_golden            # writes tests/golden/*.npz and layout.json

TEST INFRASTRUCTURE.  Each vector set is produced by the reference nn.Modules (imported from
/root/reference through `oracle/ref_import.py`) on CPU fp32 with the deterministic weights of
`oracle/synth.py`; the same run asserts that `oracle/headct_oracle.py` reproduces them, which
is what pins the oracle.  HU windowing (a1) is the one row that cannot be pinned this way: it
is MONAI's `ScaleIntensityRange` (monai==1.3.2 per setup.py:13; not installed, not vendored),
so its golden is the published MONAI formula restated -- "parity unpinned" for a1 only.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import headct_oracle as O  # noqa: E402
from oracle import ref_import, synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SMALL_GRADS = ("cls_token", "mask_token", "decoder_cls_token", "norm.weight", "norm.bias",
               "decoder_norm.weight", "blocks.0.attn.qkv.bias", "blocks.1.mlp.linear2.bias",
               "decoder_blocks.0.att_norm.weight", "decoder_pred.bias",
               "patch_embedding.patch_embeddings.bias")


def _close(a, b, tol, what):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)
    assert err <= tol, f"oracle != reference for {what}: rel err {err:.3e} > {tol}"
    return err


def _layout(model):
    return [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items()], \
           [[k, bool(p.requires_grad)] for k, p in model.named_parameters()]


def mae_case(ns, cfg, batch, x_seed, noise_seed, w_seed, name, full_outputs):
    torch.manual_seed(0)
    model = ns.mae.MaskedAutoencoderViT(**cfg)
    sd = synth.mae_state_dict(cfg, seed=w_seed)
    assert list(model.state_dict().keys()) == list(sd.keys()), "synth key order != reference"
    model.load_state_dict(sd, strict=True)
    model.train()
    x = synth.volume(batch, cfg["in_chans"], cfg["input_size"], x_seed)
    L = (cfg["input_size"] // cfg["patch_size"]) ** 3

    torch.manual_seed(noise_seed)
    noise = torch.rand(batch, L)
    assert all(len(torch.unique(r)) == L for r in noise), "pick a tie-free noise seed"

    # staged run (same generator state before each call that draws noise)
    torch.manual_seed(noise_seed)
    latent, mask, ids_restore = model.forward_encoder(x)
    pred = model.forward_decoder(latent, ids_restore)
    loss_staged = model.forward_loss(x, pred, mask)
    torch.manual_seed(noise_seed)
    tokens = model.patch_embedding(x)
    _, mask2, ids_restore2, ids_keep = model.random_masking(tokens)
    assert torch.equal(ids_restore, ids_restore2) and torch.equal(mask, mask2)

    # full forward + backward
    model.zero_grad()
    torch.manual_seed(noise_seed)
    loss, _, _ = model(x)
    loss.backward()
    assert abs(loss.item() - loss_staged.item()) < 1e-6
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}

    # ---- pin the oracle ----
    sdg = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    out = O.mae_forward(sdg, x, noise, patch=(cfg["patch_size"],) * 3, mask_ratio=cfg["mask_ratio"],
                        enc_heads=cfg["encoder_num_heads"], dec_heads=cfg["decoder_num_heads"],
                        norm_pix=cfg["norm_pix_loss"])
    assert torch.equal(out["ids_restore"], ids_restore), "ids_restore mismatch"
    assert torch.equal(out["ids_keep"], ids_keep), "ids_keep mismatch"
    assert torch.equal(out["mask"], mask), "mask mismatch"
    e1 = _close(out["latent"].detach(), latent.detach(), 2e-4, name + ".latent")
    e2 = _close(out["pred"].detach(), pred.detach(), 2e-4, name + ".pred")
    e3 = _close(out["loss"].detach(), loss.detach(), 1e-5, name + ".loss")
    out["loss"].backward()
    eg = 0.0
    for k, gref in grads.items():
        eg = max(eg, _close(sdg[k].grad, gref, 2e-3, name + ".grad." + k))
    print(f"[{name}] loss={loss.item():.6f} oracle rel err latent {e1:.1e} pred {e2:.1e} "
          f"loss {e3:.1e} grads {eg:.1e}")

    rec = dict(cfg=json.dumps(cfg), batch=batch, x_seed=x_seed, noise_seed=noise_seed, w_seed=w_seed,
               noise=noise.numpy(), loss=np.float64(loss.item()), mask=mask.numpy(),
               ids_restore=ids_restore.numpy(), ids_keep=ids_keep.numpy(),
               latent_norms=latent.detach().norm(dim=-1).numpy(),
               pred_norms=pred.detach().norm(dim=-1).numpy(),
               grad_names=np.array(list(grads.keys())),
               grad_norms=np.array([grads[k].norm().item() for k in grads], dtype=np.float64))
    if full_outputs:
        rec["latent"] = latent.detach().numpy()
        rec["pred"] = pred.detach().numpy()
        rec["tokens"] = tokens.detach().numpy()
    else:
        rec["latent_slice"] = latent.detach()[:, :4, :64].numpy()
        rec["pred_slice"] = pred.detach()[:, :4, :128].numpy()
        rec["tokens_slice"] = tokens.detach()[:, :4, :64].numpy()
    for k in SMALL_GRADS:
        if k in grads:
            rec["grad::" + k] = grads[k].numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    return model


def vit_case(ns, cfg, batch, x_seed, w_seed, name, full_outputs):
    torch.manual_seed(0)
    model = ns.vit.ViT(**synth.resolve_norm(cfg, ns.layers.RMSNorm))
    sd = synth.vit_state_dict(cfg, seed=w_seed)
    assert list(model.state_dict().keys()) == list(sd.keys()), (list(model.state_dict().keys())[:6], list(sd.keys())[:6])
    model.load_state_dict(sd, strict=True)
    model.eval()
    x = synth.volume(batch, cfg["in_chans"], cfg["img_size"], x_seed)
    with torch.no_grad():
        y, hidden = model(x)
        yo, ho = O.vit_forward(sd, x, cfg["num_heads"])
    e1 = _close(yo, y, 2e-4, name + ".tokens")
    e2 = _close(ho[-1], hidden[-1], 2e-4, name + ".hidden[-1]")
    nreg = cfg.get("num_register_tokens", 0)
    rec = dict(cfg=json.dumps(cfg), batch=batch, x_seed=x_seed, w_seed=w_seed,
               cls=y[:, 0].numpy(), pooled=y[:, 1 + nreg:].mean(1).numpy(),
               token_norms=y.norm(dim=-1).numpy(),
               hidden_norms=np.stack([h.norm(dim=-1).numpy() for h in hidden]))
    if full_outputs:
        rec["tokens"] = y.numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    print(f"[{name}] out {tuple(y.shape)} oracle rel err {e1:.1e} / {e2:.1e}")
    return model


def dino_case(ns, vit_cfg, head_cfg, batch, name):
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("gloo", rank=0, world_size=1)
    torch.manual_seed(0)
    hkw = dict(in_dim=head_cfg["in_dim"], out_dim=head_cfg["out_dim"], nlayers=head_cfg["nlayers"],
               hidden_dim=head_cfg["hidden_dim"], bottleneck_dim=head_cfg["bottleneck_dim"])
    student = ns.misc.MultiCropWrapper(ns.vit.ViT(**vit_cfg), ns.dino_head.DINOHead(**hkw))
    teacher = ns.misc.MultiCropWrapper(ns.vit.ViT(**vit_cfg), ns.dino_head.DINOHead(**hkw))
    sd_s = {**{"backbone." + k: v for k, v in synth.vit_state_dict(vit_cfg, seed=11).items()},
            **{"head." + k: v for k, v in synth.dino_head_state_dict(head_cfg, seed=12).items()}}
    sd_t = {**{"backbone." + k: v for k, v in synth.vit_state_dict(vit_cfg, seed=13).items()},
            **{"head." + k: v for k, v in synth.dino_head_state_dict(head_cfg, seed=14).items()}}
    assert list(student.state_dict().keys()) == list(sd_s.keys())
    student.load_state_dict(sd_s, strict=True)
    teacher.load_state_dict(sd_t, strict=True)
    for p in teacher.parameters():
        p.requires_grad = False
    crops = [synth.volume(batch, vit_cfg["in_chans"], vit_cfg["img_size"], 100 + i) for i in range(4)]
    crit = ns.losses.DINOLoss(head_cfg["out_dim"], 4, 0.04, 0.04, 30, 200)
    crit.center.copy_(torch.from_numpy(np.random.default_rng(5).standard_normal(
        (1, head_cfg["out_dim"])).astype(np.float32)) * 0.01)
    center0 = crit.center.clone()
    t_out = teacher(crops[:2])["dino_output"]
    s_out = student(crops)["dino_output"]
    loss = crit(s_out, t_out, 0)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in student.named_parameters() if p.grad is not None}
    # EMA (misc.py:386-397)
    m = 0.996
    ns.misc._update_momentum_encoder(student, teacher, m)

    # ---- pin the oracle ----
    bs = {k[len("backbone."):]: v for k, v in sd_s.items() if k.startswith("backbone.")}
    hs = {k[len("head."):]: v for k, v in sd_s.items() if k.startswith("head.")}
    bt = {k[len("backbone."):]: v for k, v in sd_t.items() if k.startswith("backbone.")}
    ht = {k[len("head."):]: v for k, v in sd_t.items() if k.startswith("head.")}
    with torch.no_grad():
        so = O.multicrop_forward(bs, hs, crops, vit_cfg["num_heads"])
        to = O.multicrop_forward(bt, ht, crops[:2], vit_cfg["num_heads"])
        lo = O.dino_loss(so, to, center0, ncrops=4, teacher_temp=0.04)
        co = O.dino_center_update(center0, to)
    e = [_close(so, s_out.detach(), 3e-4, name + ".student"), _close(to, t_out.detach(), 3e-4, name + ".teacher"),
         _close(lo, loss.detach(), 1e-4, name + ".loss"), _close(co, crit.center, 1e-4, name + ".center")]
    tpar = [v.clone() for k, v in sd_t.items() if k != "head.last_layer.weight_g" or True]
    # EMA oracle over parameters() order == state_dict order here (no buffers in these modules)
    t_list = [sd_t[k].clone() for k, _ in teacher.named_parameters()]
    s_list = [sd_s[k] for k, _ in student.named_parameters()]
    O.ema_update(t_list, s_list, m)
    for (k, p), t in zip(teacher.named_parameters(), t_list):
        _close(t, p.detach(), 1e-6, name + ".ema." + k)
    print(f"[{name}] loss={loss.item():.6f} oracle rel errs {['%.1e' % v for v in e]}")
    rec = dict(vit_cfg=json.dumps(vit_cfg), head_cfg=json.dumps(head_cfg), batch=batch,
               center0=center0.numpy(), loss=np.float64(loss.item()), center1=crit.center.numpy(),
               student_slice=s_out.detach()[:, :256].numpy(), teacher_slice=t_out.detach()[:, :256].numpy(),
               student_rowsum=s_out.detach().sum(1).numpy(), ema_momentum=m,
               ema_cls_token=dict(teacher.named_parameters())["backbone.cls_token"].detach().numpy(),
               grad_names=np.array(list(grads.keys())),
               grad_norms=np.array([grads[k].norm().item() for k in grads], dtype=np.float64))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    return student


def dino_head_full_case(ns, name):
    """Full-size head (768 -> 65536) + loss on random CLS features (backbone excluded)."""
    import torch.distributed as dist
    cfg = synth.DINO_HEAD_FULL
    torch.manual_seed(0)
    head_s = ns.dino_head.DINOHead(**{k: cfg[k] for k in ("in_dim", "out_dim", "nlayers", "hidden_dim", "bottleneck_dim")})
    head_t = ns.dino_head.DINOHead(**{k: cfg[k] for k in ("in_dim", "out_dim", "nlayers", "hidden_dim", "bottleneck_dim")})
    sd_s, sd_t = synth.dino_head_state_dict(cfg, seed=21), synth.dino_head_state_dict(cfg, seed=22)
    assert list(head_s.state_dict().keys()) == list(sd_s.keys())
    head_s.load_state_dict(sd_s); head_t.load_state_dict(sd_t)
    B = 2
    rng = np.random.default_rng(23)
    cls_s = torch.from_numpy(rng.standard_normal((4 * B, 768)).astype(np.float32))
    cls_t = torch.from_numpy(rng.standard_normal((2 * B, 768)).astype(np.float32))
    crit = ns.losses.DINOLoss(cfg["out_dim"], 4, 0.04, 0.04, 30, 200)
    s_out = head_s(cls_s)
    with torch.no_grad():
        t_out = head_t(cls_t)
    loss = crit(s_out, t_out, 3)
    with torch.no_grad():
        so, to = O.dino_head(sd_s, cls_s), O.dino_head(sd_t, cls_t)
        lo = O.dino_loss(so, to, torch.zeros(1, cfg["out_dim"]), ncrops=4, teacher_temp=0.04)
    _close(so, s_out.detach(), 2e-4, name + ".student"); _close(lo, loss.detach(), 1e-5, name + ".loss")
    print(f"[{name}] loss={loss.item():.6f}  (ln 65536 = {np.log(65536):.4f})")
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), batch=B, loss=np.float64(loss.item()),
                        student_slice=s_out.detach()[:, :128].numpy(), teacher_slice=t_out[:, :128].numpy(),
                        center1_slice=crit.center[:, :512].numpy(),
                        center1_sum=np.float64(crit.center.double().sum().item()))


def attention_classifier_case(ns, name):
    """classifier.py:35-100 in train mode (batch statistics) and eval mode (running statistics), 1 and 3 queries."""
    rec = {}
    for tag, dim, heads, nq, bias, B, N in (("q1", 192, 3, 1, False, 4, 69), ("q3", 96, 2, 3, True, 3, 33)):
        torch.manual_seed(0)
        clf = ns.classifier.AttentionClassifier(dim, 2, num_heads=heads, qkv_bias=bias, num_queries=nq)
        sd = synth.attention_classifier_state_dict(dim, 2, num_queries=nq, qkv_bias=bias, seed=51)
        assert list(clf.state_dict().keys()) == list(sd.keys()), list(clf.state_dict().keys())
        clf.load_state_dict(sd)
        clf.train()
        x = torch.from_numpy(np.random.default_rng(52).standard_normal((B, N, dim)).astype(np.float32)) * 1.5 + 0.3
        xg = x.clone().requires_grad_(True)
        logits = clf(xg)
        # per-sample weights: a batch-constant weighting has zero gradient through train-mode bn2 (its outputs sum to 0)
        wl = torch.from_numpy(np.random.default_rng(53).standard_normal((B, 2)).astype(np.float32))
        (logits * wl).sum().backward()
        _close(O.attention_classifier(sd, x, heads, training=True), logits.detach(), 1e-5, name + tag + ".train")
        sd_after = {k: v.clone() for k, v in clf.state_dict().items()}
        clf.eval()
        with torch.no_grad():
            logits_eval = clf(x)
        _close(O.attention_classifier(sd_after, x, heads, training=False), logits_eval, 1e-5, name + tag + ".eval")
        rec.update({f"{tag}_x": x.numpy(), f"{tag}_logits_train": logits.detach().numpy(), f"{tag}_logits_eval": logits_eval.numpy(),
                    f"{tag}_wl": wl.numpy(), f"{tag}_dx": xg.grad.numpy(), f"{tag}_dcls": clf.cls_token.grad.numpy(), f"{tag}_dwkv": clf.wkv.weight.grad.numpy(),
                    f"{tag}_bn1_mean": sd_after["bn1.running_mean"].numpy(), f"{tag}_bn1_var": sd_after["bn1.running_var"].numpy(),
                    f"{tag}_cfg": json.dumps(dict(dim=dim, heads=heads, nq=nq, bias=bias))})
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    print(f"[{name}] written")
    return clf


def lora_grad_case(ns, name):
    """LoRA fine-tuning (TRAIN.LORA, misc.py:349-359): gradients of the trainable subset through the reshape quirk."""
    cfg = synth.VIT_SMALL_LORA
    torch.manual_seed(0)
    model = ns.vit.ViT(**cfg)
    sd = synth.vit_state_dict(cfg, seed=6)
    model.load_state_dict(sd, strict=True)
    ns.misc.set_requires_grad_false(model, lora=True)
    model.train()
    x = synth.volume(2, cfg["in_chans"], cfg["img_size"], 5)
    y, _ = model(x)
    w = torch.from_numpy(np.random.default_rng(61).standard_normal(tuple(y.shape)).astype(np.float32))
    (y * w).sum().backward()
    names = ["blocks.0.attn.lora_q.lora_matrix_A", "blocks.0.attn.lora_q.lora_matrix_B", "blocks.1.attn.lora_v.lora_matrix_A",
             "blocks.1.attn.lora_v.lora_matrix_B", "blocks.0.attn.qkv.bias", "blocks.1.att_norm.weight", "norm.bias",
             "patch_embedding.patch_embeddings.bias"]
    params = dict(model.named_parameters())
    rec = {"grad::" + n: params[n].grad.numpy() for n in names}
    rec["trainable"] = json.dumps(sorted(n for n, p in params.items() if p.requires_grad))
    rec["frozen_have_no_grad"] = np.bool_(all(p.grad is None for p in params.values() if not p.requires_grad))
    rec["out_weight_seed"] = 61
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    print(f"[{name}] {len(names)} gradients written")


def misc_cases(ns):
    # a3: sin-cos table (cubic full size + a non-cubic grid that exposes the h/w swap)
    t_full = ns.pos_embed.build_sincos_position_embedding((8, 8, 8), 768, 3).detach()
    t_odd = ns.pos_embed.build_sincos_position_embedding((2, 3, 4), 12, 3).detach()
    _close(O.sincos_pos_embed_3d((8, 8, 8), 768), t_full, 1e-6, "sincos.full")
    _close(O.sincos_pos_embed_3d((2, 3, 4), 12), t_odd, 1e-6, "sincos.odd")
    # a16: linear classifier (train mode = batch statistics)
    torch.manual_seed(0)
    clf = ns.classifier.LinearClassifier(768, 2)
    sd = synth.linear_classifier_state_dict(768, 2, seed=31)
    clf.load_state_dict(sd); clf.train()
    feats = torch.from_numpy(np.random.default_rng(32).standard_normal((16, 768)).astype(np.float32))
    logits = clf(feats).detach()
    _close(O.linear_classifier(sd, feats, training=True), logits, 1e-5, "linear_classifier")
    # a1: windowing -- MONAI formula restated (unpinned, see module docstring)
    hu = synth.hu_volume(1, 8, 41)
    win = O.window_scale_stack(hu)
    np.savez_compressed(os.path.join(GOLD, "misc.npz"),
                        sincos_full_rows=t_full[0, ::37].numpy(), sincos_full_sum=np.float64(t_full.double().sum().item()),
                        sincos_full_abs_sum=np.float64(t_full.double().abs().sum().item()),
                        sincos_odd=t_odd.numpy(), clf_feats=feats.numpy(), clf_logits=logits.numpy(),
                        window_hu=hu.numpy(), window_out=win.numpy())
    print("[misc] sincos / classifier / windowing written")


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    ns = ref_import.load()
    if "--only-attention-classifier" in sys.argv:
        attention_classifier_case(ns, "attention_classifier")
        return
    layout = {}
    m = mae_case(ns, synth.MAE_SMALL, 3, 1, 7, 2, "mae_small", True)
    m = mae_case(ns, synth.MAE_FULL, 2, 3, 42, 4, "mae_full_b2", False)
    layout["mae_full"] = _layout(m)
    v = vit_case(ns, synth.VIT_SMALL, 3, 5, 6, "vit_small", True)
    v = vit_case(ns, synth.VIT_FULL_EXTRACT, 2, 7, 8, "vit_full_extract_b2", False)
    layout["vit_full_extract"] = _layout(v)
    v = vit_case(ns, synth.VIT_FULL_DINO, 1, 9, 10, "vit_full_dino_b1", False)
    layout["vit_full_dino"] = _layout(v)
    v = vit_case(ns, synth.VIT_SMALL_LORA, 2, 5, 6, "vit_small_lora", True)
    layout["vit_small_lora"] = _layout(v)
    v = vit_case(ns, synth.VIT_SMALL_RMS, 2, 5, 6, "vit_small_rms", True)
    layout["vit_small_rms"] = _layout(v)
    lora_grad_case(ns, "vit_small_lora_grads")
    c = attention_classifier_case(ns, "attention_classifier")
    layout["attention_classifier"] = _layout(c)
    s = dino_case(ns, synth.VIT_SMALL, synth.DINO_HEAD_SMALL, 2, "dino_small")
    layout["dino_small_wrapper"] = _layout(s)
    dino_head_full_case(ns, "dino_head_full")
    layout["linear_classifier"] = _layout(ns.classifier.LinearClassifier(768, 2))
    misc_cases(ns)
    # default-initialisation fingerprints: same seed -> same init as the reference constructors
    init = {}
    for nm, ctor, cfg in (("mae_small", ns.mae.MaskedAutoencoderViT, synth.MAE_SMALL),
                          ("vit_small", ns.vit.ViT, synth.VIT_SMALL)):
        torch.manual_seed(123)
        mod = ctor(**cfg)
        init[nm] = {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in mod.state_dict().items()}
    torch.manual_seed(123)
    hd = ns.dino_head.DINOHead(**synth.DINO_HEAD_SMALL)
    init["dino_head_small"] = {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in hd.state_dict().items()}
    layout["init_fingerprints_seed123"] = init
    with open(os.path.join(GOLD, "layout.json"), "w") as f:
        json.dump(layout, f, indent=0)
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    main()
