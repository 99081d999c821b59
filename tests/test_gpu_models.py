"""End-to-end GPU parity: the drop-in modules vs (a) golden vectors produced by the unmodified reference and
(b) the CPU oracle executed live on the same seeded inputs.  Tolerances are north_star's: indices bit-exact,
loss within 1e-2 relative (bf16), features cosine >= 0.999."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm())).item()


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _mae(cfg, w_seed, cuda):
    import headct_foundation_b200 as H
    from oracle import synth
    sd = synth.mae_state_dict(cfg, seed=w_seed)
    m = H.MaskedAutoencoderViT(**cfg)
    m.load_state_dict(sd, strict=True)
    return m.to(cuda).train(), sd


@pytest.mark.parametrize("name", ["mae_small", "mae_full_b2"])
def test_mae_against_golden_and_oracle(cuda, name):
    from oracle import headct_oracle as O, synth
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = json.loads(str(gold["cfg"]))
    model, sd = _mae(cfg, int(gold["w_seed"]), cuda)
    x = synth.volume(int(gold["batch"]), cfg["in_chans"], cfg["input_size"], int(gold["x_seed"]))
    noise = torch.from_numpy(gold["noise"])
    model.noise_override = noise.to(cuda)
    xc = x.to(cuda)

    latent, mask, ids_restore = model.forward_encoder(xc)
    assert torch.equal(ids_restore.cpu(), torch.from_numpy(gold["ids_restore"]))      # bit exact
    assert torch.equal(mask.cpu(), torch.from_numpy(gold["mask"]))
    pred = model.forward_decoder(latent, ids_restore)
    loss_staged = model.forward_loss(xc, pred, mask)
    _, _, _, ids_keep = model.random_masking(torch.zeros(x.shape[0], noise.shape[1], 8, device=cuda))
    assert torch.equal(ids_keep.cpu(), torch.from_numpy(gold["ids_keep"]))

    assert _rel(latent.norm(dim=-1).cpu(), torch.from_numpy(gold["latent_norms"])) < 1e-2
    assert _rel(pred.float().norm(dim=-1).cpu(), torch.from_numpy(gold["pred_norms"])) < 1e-2
    if "latent" in gold.files:
        assert _cos(latent.cpu(), torch.from_numpy(gold["latent"])) > 0.9995
        assert _cos(pred.float().cpu(), torch.from_numpy(gold["pred"])) > 0.9995
    else:
        assert _cos(latent[:, :4, :64].cpu(), torch.from_numpy(gold["latent_slice"])) > 0.999
        assert _cos(pred[:, :4, :128].float().cpu(), torch.from_numpy(gold["pred_slice"])) > 0.999

    model.zero_grad()
    loss, a, b = model(xc)
    assert a is None and b is None
    loss.backward()
    gl = float(gold["loss"])
    assert abs(loss.item() - gl) / gl < 1e-2, (loss.item(), gl)
    assert abs(loss_staged.item() - gl) / gl < 1e-2

    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    names = [str(n) for n in gold["grad_names"]]
    assert set(names) == set(grads.keys())
    gn = dict(zip(names, gold["grad_norms"]))
    bad = {k: (grads[k].norm().item(), gn[k]) for k in names if abs(grads[k].norm().item() - gn[k]) > 0.05 * gn[k] + 1e-7}
    assert not bad, bad
    for k in gold.files:
        if k.startswith("grad::"):
            assert _cos(grads[k[6:]].cpu(), torch.from_numpy(gold[k])) > 0.99, k

    # live oracle on the same inputs (independent of the stored vectors)
    if name == "mae_small":
        out = O.mae_forward(sd, x, noise, patch=(cfg["patch_size"],) * 3, mask_ratio=cfg["mask_ratio"],
                            enc_heads=cfg["encoder_num_heads"], dec_heads=cfg["decoder_num_heads"],
                            norm_pix=cfg["norm_pix_loss"])
        assert abs(loss.item() - out["loss"].item()) / out["loss"].item() < 1e-2
        assert torch.equal(ids_restore.cpu(), out["ids_restore"])


def test_mae_state_dict_roundtrip_and_default_noise(cuda):
    import headct_foundation_b200 as H
    from oracle import synth
    cfg = synth.MAE_SMALL
    m, sd = _mae(cfg, 2, cuda)
    back = m.state_dict()
    assert list(back.keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(back[k].cpu(), sd[k]), k
    # default path draws its own noise from the CUDA generator, reproducibly
    x = synth.volume(2, 3, 48, 1).to(cuda)
    torch.manual_seed(5); l1 = m(x)[0].item()
    torch.manual_seed(5); l2 = m(x)[0].item()
    torch.manual_seed(6); l3 = m(x)[0].item()
    assert l1 == l2 and l1 != l3
    with torch.no_grad():
        torch.manual_seed(5)
        assert abs(m(x)[0].item() - l1) < 1e-6


@pytest.mark.parametrize("name", ["vit_small", "vit_full_extract_b2", "vit_full_dino_b1"])
def test_vit_features(cuda, name):
    import headct_foundation_b200 as H
    from oracle import synth
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = json.loads(str(gold["cfg"]))
    sd = synth.vit_state_dict(cfg, seed=int(gold["w_seed"]))
    m = H.ViT(**cfg)
    m.load_state_dict(sd, strict=True)
    m = m.to(cuda).eval()
    x = synth.volume(int(gold["batch"]), cfg["in_chans"], cfg["img_size"], int(gold["x_seed"])).to(cuda)
    with torch.no_grad():
        y, hidden = m(x)
    nreg = cfg.get("num_register_tokens", 0)
    assert y.shape[1] == 1 + nreg + (cfg["img_size"] // cfg["patch_size"]) ** 3 and len(hidden) == cfg["num_layers"]
    for b in range(y.shape[0]):
        assert _cos(y[b, 0].cpu(), torch.from_numpy(gold["cls"][b])) >= 0.999
        assert _cos(y[b, 1 + nreg:].mean(0).cpu(), torch.from_numpy(gold["pooled"][b])) >= 0.999
    assert _rel(y.norm(dim=-1).cpu(), torch.from_numpy(gold["token_norms"])) < 1e-2
    assert _rel(hidden[-1].norm(dim=-1).cpu(), torch.from_numpy(gold["hidden_norms"][-1])) < 1e-2
    if "tokens" in gold.files:
        assert _cos(y.cpu(), torch.from_numpy(gold["tokens"])) > 0.9995


def test_vit_backward_matches_oracle(cuda):
    import headct_foundation_b200 as H
    from oracle import headct_oracle as O, synth
    cfg = synth.VIT_SMALL
    sd = synth.vit_state_dict(cfg, seed=6)
    m = H.ViT(**cfg); m.load_state_dict(sd); m = m.to(cuda).train()
    x = synth.volume(2, 3, 48, 5)
    w = torch.from_numpy(np.random.default_rng(1).standard_normal((2, 69, 192)).astype(np.float32))
    y, hidden = m(x.to(cuda))
    ((y * w.to(cuda)).sum() + hidden[0].sum() * 0.01).backward()
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    yo, ho = O.vit_forward(sdg, x, cfg["num_heads"])
    ((yo * w).sum() + ho[0].sum() * 0.01).backward()
    for k, p in m.named_parameters():
        ref = sdg[k].grad
        assert ref is not None, k
        assert _cos(p.grad.cpu(), ref) > 0.99, (k, _cos(p.grad.cpu(), ref))
        assert abs(p.grad.norm().item() - ref.norm().item()) < 0.05 * ref.norm().item() + 1e-6, k


def test_dino_step_small(cuda):
    import headct_foundation_b200 as H
    from oracle import synth
    gold = np.load(os.path.join(GOLD, "dino_small.npz"))
    vcfg, hcfg, B = json.loads(str(gold["vit_cfg"])), json.loads(str(gold["head_cfg"])), int(gold["batch"])
    def build(s1, s2):
        w = H.MultiCropWrapper(H.ViT(**vcfg), H.DINOHead(**hcfg))
        sd = {**{"backbone." + k: v for k, v in synth.vit_state_dict(vcfg, seed=s1).items()},
              **{"head." + k: v for k, v in synth.dino_head_state_dict(hcfg, seed=s2).items()}}
        w.load_state_dict(sd, strict=True)
        return w.to(cuda).train()
    student, teacher = build(11, 12), build(13, 14)
    for p in teacher.parameters():
        p.requires_grad = False
    crops = [synth.volume(B, vcfg["in_chans"], vcfg["img_size"], 100 + i).to(cuda) for i in range(4)]
    crit = H.DINOLoss(hcfg["out_dim"], 4, 0.04, 0.04, 30, 200).to(cuda)
    crit.center.copy_(torch.from_numpy(gold["center0"]).to(cuda))
    with torch.no_grad():
        t_out = teacher(crops[:2])["dino_output"]
    s_out = student(crops)["dino_output"]
    assert s_out.shape == (4 * B, hcfg["out_dim"]) and t_out.shape == (2 * B, hcfg["out_dim"])
    assert _cos(s_out[:, :256].detach().cpu(), torch.from_numpy(gold["student_slice"])) > 0.999
    assert _cos(t_out[:, :256].cpu(), torch.from_numpy(gold["teacher_slice"])) > 0.999
    loss = crit(s_out, t_out, 0)
    gl = float(gold["loss"])
    assert abs(loss.item() - gl) / gl < 1e-2, (loss.item(), gl)
    assert _rel(crit.center.cpu(), torch.from_numpy(gold["center1"])) < 2e-2
    loss.backward()
    grads = {k: p.grad for k, p in student.named_parameters() if p.grad is not None}
    names = [str(n) for n in gold["grad_names"]]
    assert set(names) == set(grads.keys())
    gn = dict(zip(names, gold["grad_norms"]))
    bad = {k: (grads[k].norm().item(), gn[k]) for k in names if abs(grads[k].norm().item() - gn[k]) > 0.08 * gn[k] + 1e-7}
    assert not bad, bad
    H.update_momentum_encoder(student, teacher, float(gold["ema_momentum"]))
    got = dict(teacher.named_parameters())["backbone.cls_token"].detach().cpu()
    assert _rel(got, torch.from_numpy(gold["ema_cls_token"])) < 1e-6
    # the EMA launch rewrites the teacher's bf16 GEMM copies: a strongly moved teacher (m = 0.5) must give the output
    # of a freshly built module holding the same fp32 weights, not the output of its previous weights
    H.update_momentum_encoder(student, teacher, 0.5)
    with torch.no_grad():
        t_after = teacher(crops[:2])["dino_output"]
    fresh = H.MultiCropWrapper(H.ViT(**vcfg), H.DINOHead(**hcfg)).to(cuda).train()
    fresh.load_state_dict(teacher.state_dict(), strict=True)
    with torch.no_grad():
        t_fresh = fresh(crops[:2])["dino_output"]
    assert torch.equal(t_after, t_fresh)
    assert _rel(t_after, t_out) > 1e-3


def test_dino_head_full_size(cuda):
    import headct_foundation_b200 as H
    from oracle import synth
    gold = np.load(os.path.join(GOLD, "dino_head_full.npz"))
    cfg = synth.DINO_HEAD_FULL
    B = int(gold["batch"])
    hs, ht = H.DINOHead(**cfg), H.DINOHead(**cfg)
    hs.load_state_dict(synth.dino_head_state_dict(cfg, seed=21)); ht.load_state_dict(synth.dino_head_state_dict(cfg, seed=22))
    hs, ht = hs.to(cuda), ht.to(cuda)
    rng = np.random.default_rng(23)
    cls_s = torch.from_numpy(rng.standard_normal((4 * B, 768)).astype(np.float32)).to(cuda)
    cls_t = torch.from_numpy(rng.standard_normal((2 * B, 768)).astype(np.float32)).to(cuda)
    crit = H.DINOLoss(cfg["out_dim"], 4, 0.04, 0.04, 30, 200).to(cuda)
    s_out = hs(cls_s)
    with torch.no_grad():
        t_out = ht(cls_t)
    loss = crit(s_out, t_out, 3)
    gl = float(gold["loss"])
    assert abs(loss.item() - gl) / gl < 1e-2
    assert _cos(s_out[:, :128].detach().cpu(), torch.from_numpy(gold["student_slice"])) > 0.999
    assert abs(crit.center.double().sum().item() - float(gold["center1_sum"])) < 2e-2 * abs(float(gold["center1_sum"])) + 1e-3
    loss.backward()
    assert hs.last_layer.weight_v.grad is not None and hs.last_layer.weight_g.grad is None


def test_training_steps_use_updated_weights(cuda):
    """After FusedAdamW steps (parameters and their bf16 GEMM copies are rewritten through raw pointers) the next
    forward must be the forward of the UPDATED fp32 weights: compare with the oracle evaluated on state_dict()."""
    import headct_foundation_b200 as H
    from headct_foundation_b200.optim import FusedAdamW
    from oracle import headct_oracle as O, synth
    gold = np.load(os.path.join(GOLD, "mae_small.npz"))
    cfg = json.loads(str(gold["cfg"]))
    model, _ = _mae(cfg, int(gold["w_seed"]), cuda)
    x = synth.volume(4, cfg["in_chans"], cfg["input_size"], 5).to(cuda)
    L = model.patch_embedding.n_patches
    opt = FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05,
                     clip_grad=3.0)
    losses = []
    for step in range(6):
        noise = synth.noise(4, L, seed=100)              # same mask every step: the loss must go down
        model.noise_override = noise.to(cuda)
        opt.zero_grad(set_to_none=True)
        loss, _, _ = model(x)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    noise = synth.noise(4, L, seed=999)
    model.noise_override = noise.to(cuda)
    with torch.no_grad():
        got = model(x)[0].item()
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    want = O.mae_forward(sd, x.cpu(), noise, patch=(cfg["patch_size"],) * 3, mask_ratio=cfg["mask_ratio"],
                         enc_heads=cfg["encoder_num_heads"], dec_heads=cfg["decoder_num_heads"],
                         norm_pix=cfg["norm_pix_loss"])["loss"].item()
    assert abs(got - want) <= 1e-2 * abs(want), (got, want, losses)
    assert losses[-1] < losses[0]                       # and the optimisation actually moves the loss


def test_torch_compile_traces_through_the_ops(cuda):
    """main_downstream.py:162 wraps the model in torch.compile.  The kernels are registered as torch.library ops
    (headct_foundation_b200/ops.py: custom_op + register_fake + register_autograd), so the tracer goes THROUGH ViT,
    PatchEmbeddingBlock and AttentionBlock: zero graph breaks on the fine-tune step, and the compiled step gives the
    eager results bit for bit, forward and backward."""
    import headct_foundation_b200 as H
    from headct_foundation_b200 import ops
    cfg = dict(in_chans=3, img_size=(24, 24, 24), patch_size=(12, 12, 12), hidden_size=96, mlp_dim=192, num_layers=2,
               num_heads=2, pos_embed="sincos", qkv_bias=True)
    torch.manual_seed(0)
    m = H.ViT(**cfg).to(cuda).train()
    clf = H.LinearClassifier(96, 2).to(cuda)
    x = torch.rand(4, 3, 24, 24, 24, device=cuda)
    y = torch.tensor([0, 1, 1, 0], device=cuda)

    def fwd(model):
        tokens, hidden = model(x)
        return torch.nn.functional.cross_entropy(clf(tokens[:, 0]), y), tokens

    def step(model):
        for p in list(m.parameters()) + list(clf.parameters()):
            p.grad = None
        loss, tokens = fwd(model)
        loss.backward()
        return loss.item(), tokens.detach().clone(), {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}

    l0, t0, g0 = step(m)
    rep = torch._dynamo.explain(lambda: fwd(m))()
    assert rep.graph_break_count == 0, rep.break_reasons
    assert rep.graph_count == 1
    torch._dynamo.reset()
    parked = ops.pending_contexts()                     # the explain() forward above had no backward: 4 contexts stay parked
    assert parked == 4
    cm = torch.compile(m)
    for _ in range(2):                                  # second call: the cached graph
        l1, t1, g1 = step(cm)
        assert l0 == l1 and torch.equal(t0, t1)
        assert set(g0) == set(g1)
        for k in g0:       # bias / norm gradients are summed with fp32 atomics: equal up to summation order
            assert torch.allclose(g0[k], g1[k], rtol=1e-4, atol=1e-6), k
        assert torch.equal(g0["blocks.0.attn.qkv.weight"], g1["blocks.0.attn.qkv.weight"])
    assert ops.pending_contexts() == parked             # every forward context of the steps was consumed by its backward
    # inference under no_grad parks nothing
    with torch.no_grad():
        t2, _ = cm(x)
    assert torch.equal(t2, t0) and ops.pending_contexts() == parked
    # RMSNorm blocks and the ops called directly in eager mode
    out = torch.ops.headct.layernorm(t0, m.norm.weight, None, 1e-6, False, 0)[0]
    assert out.shape == t0.shape and out.dtype == torch.float32


def test_torch_compile_mae_step_still_runs(cuda):
    """The MAE wrapper (masking, decoder assembly) stays opaque to the tracer (graph breaks around those methods); the
    compiled module must keep giving the eager loss."""
    import headct_foundation_b200 as H
    from oracle import synth
    cfg = synth.MAE_SMALL
    m = H.MaskedAutoencoderViT(**cfg).to(cuda).train()
    x = synth.volume(2, 3, 48, 1).to(cuda)
    m.noise_override = synth.noise(2, 64, seed=3).to(cuda)
    l0 = m(x)[0]
    l0.backward()
    g0 = m.decoder_pred.weight.grad.clone()
    m.zero_grad()
    torch._dynamo.reset()
    l1 = torch.compile(m)(x)[0]
    l1.backward()
    assert l0.item() == l1.item() and torch.allclose(g0, m.decoder_pred.weight.grad, rtol=1e-4, atol=1e-7)


def test_reference_engine_step_with_amp_gradscaler_and_torch_adamw(cuda):
    """The reference's own training step around the drop-in model (engine_pretrain_mae.py:52-86): fp16 autocast +
    GradScaler, per-parameter clip_gradients with .item() (misc.py:374-383), stock torch.optim.AdamW.  Two steps must
    follow the fp32 oracle trained the same way on the CPU (loss within 1e-2, updated weights aligned)."""
    import headct_foundation_b200 as H
    from oracle import headct_oracle as O, synth
    gold = np.load(os.path.join(GOLD, "mae_small.npz"))
    cfg = json.loads(str(gold["cfg"]))
    model, sd0 = _mae(cfg, int(gold["w_seed"]), cuda)
    x = synth.volume(4, cfg["in_chans"], cfg["input_size"], 5)
    L = model.patch_embedding.n_patches
    kw = dict(patch=(cfg["patch_size"],) * 3, mask_ratio=cfg["mask_ratio"], enc_heads=cfg["encoder_num_heads"],
              dec_heads=cfg["decoder_num_heads"], norm_pix=cfg["norm_pix_loss"])

    def clip_gradients(params, clip):                     # misc.py:374-383, on any iterable of tensors with .grad
        for p in params:
            if p.grad is not None:
                n = p.grad.data.norm(2)
                c = clip / (n + 1e-6)
                if c < 1:
                    p.grad.data.mul_(c)

    # ---- oracle side: fp32, CPU
    ref = {k: v.clone().requires_grad_(v.is_floating_point() and k != "decoder_pos_embed") for k, v in sd0.items()}
    ref_params = [v for v in ref.values() if v.requires_grad]
    ref_opt = torch.optim.AdamW(ref_params, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05)
    # ---- drop-in side: the engine's AMP recipe
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05)
    scaler = torch.amp.GradScaler("cuda")
    xc = x.to(cuda)
    for step in range(2):
        noise = synth.noise(4, L, seed=300 + step)
        ref_opt.zero_grad()
        rl = O.mae_forward(ref, x, noise, **kw)["loss"]
        rl.backward()
        clip_gradients(ref_params, 3.0)
        ref_opt.step()

        model.noise_override = noise.to(cuda)
        opt.zero_grad()
        with torch.amp.autocast("cuda", dtype=torch.float16, enabled=True):
            loss, _, _ = model(xc)
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        clip_gradients(model.parameters(), 3.0)
        scaler.step(opt)
        scaler.update()
        assert abs(loss.item() - rl.item()) <= 1e-2 * abs(rl.item()), (step, loss.item(), rl.item())
    model.noise_override = None
    got = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    for k in ("blocks.0.attn.qkv.weight", "decoder_blocks.1.mlp.linear2.weight", "decoder_pred.bias", "cls_token",
              "patch_embedding.patch_embeddings.weight"):
        moved = (ref[k].detach() - sd0[k]).norm().item()
        assert moved > 0
        assert (got[k] - ref[k].detach()).norm().item() < 0.2 * moved + 1e-6, k      # the UPDATE itself agrees


def test_graphed_forward_equals_eager(cuda):
    """CUDA-graph replay of the extraction forward (utils/graphs.py): bit-identical to the eager launches, follows new
    inputs, and re-captures after the weights changed."""
    import headct_foundation_b200 as H
    from oracle import synth
    cfg = synth.VIT_SMALL
    m = H.ViT(**cfg)
    m.load_state_dict(synth.vit_state_dict(cfg, seed=6))
    m = m.to(cuda).eval()
    x1 = synth.volume(2, 3, 48, 5).to(cuda)
    x2 = synth.volume(2, 3, 48, 9).to(cuda)
    gf = H.GraphedForward(m, x1, clone=True)
    with torch.no_grad():
        e1, h1 = m(x1)
        e2, h2 = m(x2)
    g1, gh1 = gf(x1)
    g2, gh2 = gf(x2)
    assert torch.equal(g1, e1) and torch.equal(g2, e2) and not torch.equal(g1, g2)
    assert all(torch.equal(a, b) for a, b in zip(gh2, h2)) and len(gh1) == cfg["num_layers"]
    with torch.no_grad():
        m.blocks[0].mlp.linear1.weight.mul_(1.5)                 # in-place update bumps the version counter
        e3, _ = m(x2)
    g3, _ = gf(x2)
    assert torch.equal(g3, e3) and not torch.equal(g3, g2)
    with pytest.raises(ValueError):
        gf(x2[:1])


def test_graphed_train_step_follows_eager(cuda):
    """One CUDA-graph launch per training step (utils/graphs.py) trains like the eager step: same losses and weights over
    several steps with a moving learning rate (read from device memory at replay), and capturing leaves the model and
    the optimizer state untouched."""
    import headct_foundation_b200 as H
    from headct_foundation_b200.optim import FusedAdamW
    from oracle import synth
    cfg = synth.MAE_SMALL
    sd = synth.mae_state_dict(cfg, seed=2)
    noise = synth.noise(2, (cfg["input_size"] // cfg["patch_size"]) ** 3, seed=3).to(cuda)
    xs = [synth.volume(2, cfg["in_chans"], cfg["input_size"], 20 + i).to(cuda) for i in range(4)]

    def build():
        m = H.MaskedAutoencoderViT(**cfg)
        m.load_state_dict(sd, strict=True)
        m = m.to(cuda).train()
        m.noise_override = noise
        opt = FusedAdamW([p for p in m.parameters() if p.requires_grad], lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05,
                         clip_grad=3.0)
        return m, opt

    m1, o1 = build()
    eager = []
    for i, x in enumerate(xs):
        o1.param_groups[0]["lr"] = 1e-3 * (i + 1)
        o1.zero_grad(set_to_none=True)
        loss = m1(x)[0]
        loss.backward()
        o1.step()
        eager.append(loss.item())
    m2, o2 = build()
    step = H.GraphedTrainStep(m2, o2, xs[0])
    for k, v in m2.state_dict().items():                          # the warm-up inside the capture was rolled back
        assert torch.equal(v.cpu(), sd[k]), k
    graphed = []
    for i, x in enumerate(xs):
        o2.param_groups[0]["lr"] = 1e-3 * (i + 1)
        graphed.append(step(x).item())
    assert all(abs(a - b) < 2e-3 * abs(a) for a, b in zip(eager, graphed)), (eager, graphed)
    assert eager[0] != eager[-1]
    p0 = next(iter(o2.state))
    assert int(o2.state[p0]["step"].item()) == len(xs)
    # Adam's first steps move every element by ~lr * sign(g): elements whose gradient is at the noise level of the
    # split-K atomics differ between ANY two runs, so the weights are compared by direction, the trajectory by its losses
    ups1, ups2 = [], []
    for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        assert _cos(a.detach().cpu(), b.detach().cpu()) > 0.99, k
        ups1.append((a.detach().cpu() - sd[k]).flatten()); ups2.append((b.detach().cpu() - sd[k]).flatten())
    assert _cos(torch.cat(ups1), torch.cat(ups2)) > 0.95            # the accumulated update of the whole model
    # eager use after replays sees the updated weights (bf16 copies re-keyed by advance())
    with torch.no_grad():
        l1, l2 = m1(xs[0])[0].item(), m2(xs[0])[0].item()
    assert abs(l1 - l2) < 2e-3 * abs(l1)
