"""Pins the CPU oracle to the golden vectors generated from the unmodified reference (oracle/gen_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import headct_oracle as O
from oracle import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _close(a, b, tol):
    a = torch.as_tensor(a, dtype=torch.float64); b = torch.as_tensor(b, dtype=torch.float64)
    return (a - b).abs().max().item() <= tol * max(b.abs().max().item(), 1e-12)


@pytest.mark.parametrize("name", ["mae_small", "mae_full_b2"])
def test_mae_oracle(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = json.loads(str(g["cfg"]))
    sd = synth.mae_state_dict(cfg, seed=int(g["w_seed"]))
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = synth.volume(int(g["batch"]), cfg["in_chans"], cfg["input_size"], int(g["x_seed"]))
    noise = torch.from_numpy(g["noise"])
    out = O.mae_forward(sdg, x, noise, patch=(cfg["patch_size"],) * 3, mask_ratio=cfg["mask_ratio"],
                        enc_heads=cfg["encoder_num_heads"], dec_heads=cfg["decoder_num_heads"],
                        norm_pix=cfg["norm_pix_loss"])
    assert torch.equal(out["ids_restore"], torch.from_numpy(g["ids_restore"]))
    assert torch.equal(out["ids_keep"], torch.from_numpy(g["ids_keep"]))
    assert torch.equal(out["mask"], torch.from_numpy(g["mask"]))
    assert abs(out["loss"].item() - float(g["loss"])) < 1e-5 * float(g["loss"])
    assert _close(out["latent"].detach().norm(dim=-1), g["latent_norms"], 1e-4)
    assert _close(out["pred"].detach().norm(dim=-1), g["pred_norms"], 1e-4)
    if "pred" in g.files:
        assert _close(out["pred"].detach(), g["pred"], 5e-4) and _close(out["latent"].detach(), g["latent"], 5e-4)
    out["loss"].backward()
    norms = dict(zip([str(n) for n in g["grad_names"]], g["grad_norms"]))
    for k, n in norms.items():
        assert abs(sdg[k].grad.norm().item() - n) <= 2e-3 * n + 1e-9, k
    for k in g.files:
        if k.startswith("grad::"):
            assert _close(sdg[k[6:]].grad, g[k], 3e-3), k


@pytest.mark.parametrize("name", ["vit_small", "vit_full_extract_b2", "vit_full_dino_b1", "vit_small_lora", "vit_small_rms"])
def test_vit_oracle(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = json.loads(str(g["cfg"]))
    sd = synth.vit_state_dict(cfg, seed=int(g["w_seed"]))
    x = synth.volume(int(g["batch"]), cfg["in_chans"], cfg["img_size"], int(g["x_seed"]))
    with torch.no_grad():
        y, hidden = O.vit_forward(sd, x, cfg["num_heads"])
    nreg = cfg.get("num_register_tokens", 0)
    assert _close(y[:, 0], g["cls"], 5e-4) and _close(y[:, 1 + nreg:].mean(1), g["pooled"], 5e-4)
    assert _close(torch.stack([h.norm(dim=-1) for h in hidden]), g["hidden_norms"], 2e-4)


def test_dino_oracle():
    g = np.load(os.path.join(GOLD, "dino_small.npz"))
    vcfg, hcfg, B = json.loads(str(g["vit_cfg"])), json.loads(str(g["head_cfg"])), int(g["batch"])
    crops = [synth.volume(B, vcfg["in_chans"], vcfg["img_size"], 100 + i) for i in range(4)]
    with torch.no_grad():
        so = O.multicrop_forward(synth.vit_state_dict(vcfg, 11), synth.dino_head_state_dict(hcfg, 12), crops, vcfg["num_heads"])
        to = O.multicrop_forward(synth.vit_state_dict(vcfg, 13), synth.dino_head_state_dict(hcfg, 14), crops[:2], vcfg["num_heads"])
        c0 = torch.from_numpy(g["center0"])
        loss = O.dino_loss(so, to, c0, ncrops=4, teacher_temp=0.04)
        c1 = O.dino_center_update(c0, to)
    assert _close(so[:, :256], g["student_slice"], 5e-4) and _close(to[:, :256], g["teacher_slice"], 5e-4)
    assert abs(loss.item() - float(g["loss"])) < 1e-4 * float(g["loss"])
    assert _close(c1, g["center1"], 1e-4)
    sched = O.teacher_temp_schedule(0.04, 0.07, 30, 200)
    assert len(sched) == 200 and sched[0] == 0.04 and abs(sched[29] - 0.07) < 1e-12 and sched[-1] == 0.07


def test_misc_oracle():
    g = np.load(os.path.join(GOLD, "misc.npz"))
    t = O.sincos_pos_embed_3d((8, 8, 8), 768)
    assert _close(t[0, ::37], g["sincos_full_rows"], 1e-6)
    assert abs(t.double().sum().item() - float(g["sincos_full_sum"])) < 1e-6 * float(g["sincos_full_abs_sum"])
    assert _close(O.sincos_pos_embed_3d((2, 3, 4), 12), g["sincos_odd"], 1e-6)
    sd = synth.linear_classifier_state_dict(768, 2, seed=31)
    assert _close(O.linear_classifier(sd, torch.from_numpy(g["clf_feats"])), g["clf_logits"], 1e-5)
    w = O.window_scale_stack(torch.from_numpy(g["window_hu"]))
    assert torch.equal(w, torch.from_numpy(g["window_out"]))
    assert O.window_bounds() == [(0.0, 80.0), (-20.0, 180.0), (-800.0, 2000.0)]
    hu = torch.tensor([-2000.0, 0.0, 40.0, 80.0, 180.0, 2000.0, 4000.0]).view(1, 7, 1, 1)
    out = O.window_scale_stack(hu)
    assert out.shape == (3, 7, 1, 1) and out.min() == 0 and out.max() == 1 and out[0, 2, 0, 0] == 0.5


def test_masking_edge_cases():
    ids_keep, ids_restore, mask, keep = O.masking_indices(torch.zeros(2, 8), 0.75)       # all ties -> identity
    assert keep == 2 and torch.equal(ids_restore, torch.arange(8).expand(2, 8))
    assert torch.equal(mask[0], torch.tensor([0., 0., 1., 1., 1., 1., 1., 1.]))
    assert int(512 * (1 - 0.75)) == 128 and int(64 * (1 - 0.9)) == 6                  # float-product floor (mae.py:205)
    _, _, m0, k0 = O.masking_indices(torch.rand(1, 5), 1.0)
    assert k0 == 0 and m0.sum() == 5


def test_optimizer_helpers():
    p = torch.tensor([1.0, -2.0]); g = torch.tensor([0.5, 0.25]); m = torch.zeros(2); v = torch.zeros(2)
    ref = torch.nn.Parameter(p.clone()); ref.grad = g.clone()
    opt = torch.optim.AdamW([ref], lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    for step in (1, 2, 3):
        O.adamw_step(p, g, m, v, step, lr=1e-2, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.05)
        opt.step()
    assert torch.allclose(p, ref.detach(), atol=1e-7)
    gs = [torch.tensor([3.0, 4.0]), torch.tensor([0.1])]
    norms = O.clip_per_param(gs, 1.0)
    assert norms[0] == 5.0 and abs(gs[0].norm().item() - 1.0) < 1e-5 and gs[1].item() == pytest.approx(0.1)


def test_lora_gradients_oracle():
    """LoRA fine-tuning (reshape quirk of attentionblock.py:57-59 included): oracle autograd == reference gradients."""
    g = np.load(os.path.join(GOLD, "vit_small_lora_grads.npz"))
    cfg = synth.VIT_SMALL_LORA
    sd = {k: v.clone().requires_grad_(True) for k, v in synth.vit_state_dict(cfg, seed=6).items()}
    x = synth.volume(2, cfg["in_chans"], cfg["img_size"], 5)
    y, _ = O.vit_forward(sd, x, cfg["num_heads"])
    w = torch.from_numpy(np.random.default_rng(int(g["out_weight_seed"])).standard_normal(tuple(y.shape)).astype(np.float32))
    (y * w).sum().backward()
    for k in g.files:
        if k.startswith("grad::"):
            assert _close(sd[k[6:]].grad, g[k], 2e-3), k
    assert bool(g["frozen_have_no_grad"])


def test_attention_classifier_oracle():
    g = np.load(os.path.join(GOLD, "attention_classifier.npz"))
    for tag in ("q1", "q3"):
        c = json.loads(str(g[tag + "_cfg"]))
        sd = synth.attention_classifier_state_dict(c["dim"], 2, num_queries=c["nq"], qkv_bias=c["bias"], seed=51)
        x = torch.from_numpy(g[tag + "_x"])
        assert _close(O.attention_classifier(sd, x, c["heads"], training=True), g[tag + "_logits_train"], 1e-4)
        sd["bn1.running_mean"], sd["bn1.running_var"] = torch.from_numpy(g[tag + "_bn1_mean"]), torch.from_numpy(g[tag + "_bn1_var"])
        # bn2's running statistics after one training step: restated from the train-mode pooled features
        B, N, C = x.shape
        mu, var = x.mean((0, 1)), x.var((0, 1), unbiased=False)
        assert _close(0.9 * torch.zeros(C) + 0.1 * mu, g[tag + "_bn1_mean"], 1e-5)
        assert _close(0.9 * torch.ones(C) + 0.1 * var * (B * N) / (B * N - 1), g[tag + "_bn1_var"], 1e-5)
