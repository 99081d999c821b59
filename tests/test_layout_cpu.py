"""Drop-in surface on CPU: state_dict layout, parameter order / requires_grad and default initialisation are
identical to the reference classes (fingerprints recorded by oracle/gen_golden.py from the real modules)."""
import json
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")
LAYOUT = json.load(open(os.path.join(GOLD, "layout.json")))


def _lay(m):
    return ([[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in m.state_dict().items()],
            [[k, bool(p.requires_grad)] for k, p in m.named_parameters()])


def test_mae_layout():
    import headct_foundation_b200 as H
    from oracle import synth
    m = H.MaskedAutoencoderViT(**synth.MAE_FULL)
    sd, params = _lay(m)
    assert sd == LAYOUT["mae_full"][0] and params == LAYOUT["mae_full"][1]
    assert sum(p.numel() for p in m.parameters()) == 151108416
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 150715200
    assert m.patch_embedding.n_patches == 512 and m.patch_embedding.position_embeddings.requires_grad


@pytest.mark.parametrize("key,cfg", [("vit_full_extract", "VIT_FULL_EXTRACT"), ("vit_full_dino", "VIT_FULL_DINO")])
def test_vit_layout(key, cfg):
    import headct_foundation_b200 as H
    from oracle import synth
    m = H.ViT(**getattr(synth, cfg))
    assert list(_lay(m)) == [LAYOUT[key][0], LAYOUT[key][1]]
    if key == "vit_full_extract":
        assert sum(p.numel() for p in m.parameters()) == 89404416            # notebook cell 4


def test_multicrop_and_classifier_layout():
    import headct_foundation_b200 as H
    from oracle import synth
    w = H.MultiCropWrapper(H.ViT(**synth.VIT_SMALL), H.DINOHead(**synth.DINO_HEAD_SMALL))
    assert list(_lay(w)) == [LAYOUT["dino_small_wrapper"][0], LAYOUT["dino_small_wrapper"][1]]
    names = [k for k, _ in w.named_parameters()]
    assert any("last_layer" in n for n in names)                             # misc.py:366-371 relies on the name
    assert isinstance(w.backbone.fc, torch.nn.Identity)
    c = H.LinearClassifier(768, 2)
    assert list(_lay(c)) == [LAYOUT["linear_classifier"][0], LAYOUT["linear_classifier"][1]]


def test_downstream_variant_layouts():
    """LoRA adapters, RMSNorm blocks and the attentive probe keep the reference's state_dict layout."""
    import headct_foundation_b200 as H
    from oracle import synth
    m = H.ViT(**synth.VIT_SMALL_LORA)
    assert list(_lay(m)) == [LAYOUT["vit_small_lora"][0], LAYOUT["vit_small_lora"][1]]
    H.set_requires_grad_false(m, lora=True)                                  # misc.py:349-359
    trainable = {n for n, p in m.named_parameters() if p.requires_grad}
    assert "blocks.0.attn.lora_q.lora_matrix_A" in trainable and "blocks.0.attn.qkv.bias" in trainable
    assert "patch_embedding.position_embeddings" in trainable and "norm.weight" in trainable
    assert "blocks.0.attn.qkv.weight" not in trainable and "cls_token" not in trainable
    m = H.ViT(**synth.resolve_norm(synth.VIT_SMALL_RMS, H.RMSNorm))
    assert list(_lay(m)) == [LAYOUT["vit_small_rms"][0], LAYOUT["vit_small_rms"][1]]
    c = H.AttentionClassifier(96, 2, num_heads=2, qkv_bias=True, num_queries=3)
    assert list(_lay(c)) == [LAYOUT["attention_classifier"][0], LAYOUT["attention_classifier"][1]]
    with pytest.raises(NotImplementedError):
        H.AttentionBlock(64, 128, 2, norm_layer=torch.nn.BatchNorm1d)


@pytest.mark.parametrize("name", ["mae_small", "vit_small", "dino_head_small"])
def test_same_seed_same_init_as_reference(name):
    import headct_foundation_b200 as H
    from oracle import synth
    torch.manual_seed(123)
    if name == "mae_small":
        m = H.MaskedAutoencoderViT(**synth.MAE_SMALL)
    elif name == "vit_small":
        m = H.ViT(**synth.VIT_SMALL)
    else:
        m = H.DINOHead(**synth.DINO_HEAD_SMALL)
    fp = LAYOUT["init_fingerprints_seed123"][name]
    got = m.state_dict()
    assert list(got.keys()) == list(fp.keys())
    for k, (s, a) in fp.items():
        v = got[k].double()
        assert abs(float(v.sum()) - s) <= 1e-9 * max(1.0, abs(a)) and abs(float(v.abs().sum()) - a) <= 1e-9 * max(1.0, a), k


def test_sincos_table_matches_reference_rows():
    import numpy as np
    import headct_foundation_b200 as H
    g = np.load(os.path.join(GOLD, "misc.npz"))
    t = H.build_sincos_position_embedding((8, 8, 8), 768, 3)
    assert not t.requires_grad and t.shape == (1, 512, 768)
    assert torch.equal(t[0, ::37], torch.from_numpy(g["sincos_full_rows"]))
    assert torch.equal(H.build_sincos_position_embedding((2, 3, 4), 12, 3).detach(), torch.from_numpy(g["sincos_odd"]))
