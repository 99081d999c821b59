"""Checkpoint wire format: files interchange with the reference's save_checkpoint / load_model
(src/utils/misc.py:35-96) -- same top-level keys, wrapper prefixes removed on load, strict=False semantics, optimizer
state interchangeable with torch.optim.AdamW, position embeddings resampled like pos_embed.py:102-153."""
import logging
import os

import pytest
import torch

SMALL_VIT = dict(in_chans=3, img_size=(24, 24, 24), patch_size=(12, 12, 12), hidden_size=96, mlp_dim=192, num_layers=2,
                 num_heads=4, pos_embed="sincos", qkv_bias=True, num_register_tokens=2)


def _vit(**over):
    import headct_foundation_b200 as H
    cfg = dict(SMALL_VIT)
    cfg.update(over)
    return H.ViT(**cfg)


def test_save_load_roundtrip_and_wire_format(tmp_path):
    import headct_foundation_b200 as H
    from headct_foundation_b200.utils import checkpoint as ck
    torch.manual_seed(0)
    student = H.MultiCropWrapper(_vit(), H.DINOHead(96, 64, hidden_dim=32, bottleneck_dim=16))
    teacher = H.MultiCropWrapper(_vit(), H.DINOHead(96, 64, hidden_dim=32, bottleneck_dim=16))
    opt = torch.optim.AdamW(student.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda e: 1.0)
    path = ck.save_checkpoint(student, teacher, 7, opt, sched, filename="m.pt", best_loss=0.25, dir_add=str(tmp_path),
                              logger=logging.getLogger("t"))
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert list(raw.keys()) == ["epoch", "best_loss", "state_dict", "momentum_model_state_dict", "optimizer", "scheduler"]
    assert raw["epoch"] == 7 and raw["best_loss"] == 0.25
    assert all(k.startswith(("backbone.", "head.")) for k in raw["state_dict"])
    # a bare backbone picks its tensors out of the wrapped checkpoint (prefix removal + strict=False, misc.py:78-83)
    bare = _vit()
    got = ck.load_model(path, bare, None)
    assert got["epoch"] == 7
    for k, v in bare.state_dict().items():
        assert torch.equal(v, raw["state_dict"]["backbone." + k]), k
    # DDP / torch.compile prefixes are removed too
    wrapped = {"module._orig_mod." + k: v for k, v in raw["state_dict"].items()}
    assert set(ck.strip_wrapper_prefixes(wrapped)) == {k.replace("backbone.", "") for k in raw["state_dict"]}
    # optimizer / scheduler / epoch
    opt2 = torch.optim.AdamW(student.parameters(), lr=5.0)
    sched2 = torch.optim.lr_scheduler.LambdaLR(opt2, lambda e: 1.0)
    _, _, epoch = ck.load_optimizer(opt2, sched2, raw)
    assert epoch == 7 and opt2.param_groups[0]["lr"] == 1e-3
    assert ck.load_optimizer(opt2, sched2, {})[2] == 0
    assert ck.load_model(None, bare, None) is None


def test_reference_style_config_object_and_momentum_model(tmp_path):
    import headct_foundation_b200 as H
    from headct_foundation_b200.utils import checkpoint as ck
    torch.manual_seed(1)
    a, b = _vit(), _vit()
    opt = torch.optim.AdamW(a.parameters())
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda e: 1.0)
    path = ck.save_checkpoint(a, b, 1, opt, sched, dir_add=str(tmp_path))

    class _M:
        PRETRAINED = path
    class _Cfg:
        MODEL = _M
    c, d = _vit(), _vit()
    ck.load_model(_Cfg, c, d)
    assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), c.state_dict().values()))
    assert all(torch.equal(x, y) for x, y in zip(b.state_dict().values(), d.state_dict().values()))


def test_fused_adamw_state_interchanges_with_torch_adamw():
    from headct_foundation_b200.optim import FusedAdamW
    p = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    ref = torch.optim.AdamW(p, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05)
    for q in p:
        q.grad = torch.randn_like(q)
    ref.step()
    fused = FusedAdamW(p, lr=7.0, betas=(0.5, 0.5), weight_decay=0.0, clip_grad=3.0)
    fused.load_state_dict(ref.state_dict())
    g = fused.param_groups[0]
    assert g["lr"] == 1e-3 and tuple(g["betas"]) == (0.9, 0.95) and g["weight_decay"] == 0.05
    for q in p:
        assert torch.equal(fused.state[q]["exp_avg"], ref.state[q]["exp_avg"])
        assert float(fused.state[q]["step"]) == 1.0
    back = torch.optim.AdamW(p)
    back.load_state_dict(fused.state_dict())
    assert torch.equal(back.state[p[1]]["exp_avg_sq"], ref.state[p[1]]["exp_avg_sq"])


@pytest.mark.parametrize("spatial_dims", [3])
def test_interpolate_pos_embed_matches_trilinear_resampling(spatial_dims):
    from headct_foundation_b200.utils import checkpoint as ck
    torch.manual_seed(2)
    big = _vit(img_size=(48, 48, 48), pos_embed="learnable")            # 4^3 patches
    small = _vit(pos_embed="learnable")                                  # 2^3 patches
    sd = {k: v.clone() for k, v in small.state_dict().items()}
    src = sd["patch_embedding.position_embeddings"].clone()
    ck.interpolate_pos_embed(big, sd)
    out = sd["patch_embedding.position_embeddings"]
    assert out.shape == big.patch_embedding.position_embeddings.shape == (1, 64, 96)
    want = torch.nn.functional.interpolate(src.reshape(1, 2, 2, 2, 96).permute(0, 4, 1, 2, 3), size=(4, 4, 4),
                                           mode="trilinear", align_corners=False).permute(0, 2, 3, 4, 1).reshape(1, 64, 96)
    assert torch.allclose(out, want, atol=1e-6)
    big.load_state_dict(sd, strict=True)
    same = {k: v.clone() for k, v in small.state_dict().items()}
    ck.interpolate_pos_embed(small, same)                                # same grid: untouched
    assert torch.equal(same["patch_embedding.position_embeddings"], src)


def test_checkpoint_matches_the_reference_loader(tmp_path):
    """If the reference tree is present (this container, not the GPU box): its own load_model reads our file."""
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    ref = ref_import.load()
    from headct_foundation_b200.utils import checkpoint as ck
    torch.manual_seed(3)
    ours = _vit()
    opt = torch.optim.AdamW(ours.parameters())
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda e: 1.0)
    path = ck.save_checkpoint(ours, None, 3, opt, sched, dir_add=str(tmp_path))
    theirs = ref.vit.ViT(**SMALL_VIT)

    class _M:
        PRETRAINED = path
        NAME = "vit"
    class _Cfg:
        MODEL = _M
    ref.misc.load_model(_Cfg, theirs, None, logging.getLogger("t"))
    for (k, v), (k2, v2) in zip(ours.state_dict().items(), theirs.state_dict().items()):
        assert k == k2 and torch.equal(v, v2), k
    # and the other direction: a file written by the reference's save_checkpoint loads here
    os.makedirs(tmp_path / "ref", exist_ok=True)
    ref.misc.save_checkpoint(theirs, None, 4, opt, sched, filename="r.pt", best_loss=1.0, dir_add=str(tmp_path / "ref"),
                             logger=logging.getLogger("t"))
    again = _vit()
    got = ck.load_model(str(tmp_path / "ref" / "r.pt"), again, None)
    assert got["epoch"] == 4
    assert all(torch.equal(x, y) for x, y in zip(again.state_dict().values(), theirs.state_dict().values()))
    # pos-embed resampling agrees with the reference's interpolate_pos_embed
    big_ours, big_ref = _vit(img_size=(48, 48, 48), pos_embed="learnable"), None
    sd_a = {k: v.clone() for k, v in _vit(pos_embed="learnable").state_dict().items()}
    sd_b = {k: v.clone() for k, v in sd_a.items()}
    ck.interpolate_pos_embed(big_ours, sd_a)
    ref.pos_embed.interpolate_pos_embed(big_ours, sd_b)
    assert torch.allclose(sd_a["patch_embedding.position_embeddings"], sd_b["patch_embedding.position_embeddings"], atol=1e-6)


def test_load_model_uses_the_safe_unpickler(tmp_path):
    """misc.py:73-76 loads with torch's default (safe) unpickler + a numpy-scalar allow-list.  A file that smuggles an
    arbitrary callable must be refused unless the caller opts in with trust_pickle=True; a checkpoint that carries
    `best_loss` as a numpy scalar (what the reference writes) must load."""
    import numpy as np
    from headct_foundation_b200.utils.checkpoint import load_model

    class Boom:
        def __reduce__(self):
            return (eval, ("__import__('os').environ.__setitem__('HCT_PWNED', '1')",))

    model = _vit()
    bad = tmp_path / "bad.pt"
    torch.save({"state_dict": model.state_dict(), "extra": Boom()}, bad)
    os.environ.pop("HCT_PWNED", None)
    with pytest.raises(Exception):
        load_model(str(bad), model)
    assert "HCT_PWNED" not in os.environ
    good = tmp_path / "good.pt"
    torch.save({"state_dict": model.state_dict(), "best_loss": np.float64(0.25), "epoch": 3}, good)
    ckpt = load_model(str(good), model)
    assert float(ckpt["best_loss"]) == 0.25 and ckpt["epoch"] == 3
