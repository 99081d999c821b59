"""CPU-side checks of the boundary: the C-ABI library builds/loads, exports exactly what include/hct_b200.h
declares, and the product path refuses to run without a GPU (no silent fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "hct_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hct_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    from headct_foundation_b200 import _cabi, build
    build.build()
    L = _cabi.lib()
    assert L.hct_abi_version() == 1
    declared = _declared()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"{name} declared in hct_b200.h but not exported"
    assert set(_cabi.EXPORTED_SYMBOLS) == set(declared)
    assert L.hct_launch_count() == 0          # nothing has been launched on this GPU-less box


def test_gemm_desc_matches_c_struct_layout(tmp_path):
    """ctypes mirror == the C struct as gcc lays it out from include/hct_b200.h (sizeof + every offset)."""
    import shutil
    import subprocess
    from headct_foundation_b200._cabi import GemmDesc
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    fields = [f[0] for f in GemmDesc._fields_]
    src = tmp_path / "lay.c"
    body = "\n".join(f'  printf("{f} %zu\\n", offsetof(hct_gemm_desc, {f}));' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "hct_b200.h"\nint main(void) {\n'
                   '  printf("sizeof %zu\\n", sizeof(hct_gemm_desc));\n' + body + "\n  return 0;\n}\n")
    exe = tmp_path / "lay"
    subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert int(out["sizeof"]) == ctypes.sizeof(GemmDesc)
    for f in fields:
        assert int(out[f]) == getattr(GemmDesc, f).offset, f


def test_argument_validation_without_gpu():
    from headct_foundation_b200 import _cabi
    L = _cabi.lib()
    d = _cabi.GemmDesc()
    d.M, d.N, d.K = 128, 100, 64            # N not a multiple of 8 -> rejected before any CUDA call
    assert L.hct_gemm_bf16(ctypes.byref(d), None) == 1
    assert b"multiple of 8" in L.hct_last_error()
    assert L.hct_layernorm_fwd(None, None, None, None, 1, None, None, 4, 7, 1e-5, None) == 1
    assert L.hct_attention_fwd(None, None, None, 1, 16, 2, 40, None) == 3       # unsupported head dim
    assert L.hct_mask_indices(None, None, None, None, 0, 512, 128, None) == 0    # empty batch is a no-op


def test_no_cpu_fallback():
    import headct_foundation_b200 as H
    from oracle import synth
    m = H.MaskedAutoencoderViT(**synth.MAE_SMALL)
    with pytest.raises(RuntimeError, match="CUDA-only|no CPU fallback|cuda"):
        m(torch.zeros(1, 3, 48, 48, 48))
    with pytest.raises(RuntimeError):
        H.ViT(**synth.VIT_SMALL)(torch.zeros(1, 3, 48, 48, 48))


def test_no_cpu_fallback_downstream_heads_and_graphs():
    """The rows added after the core path (8(f) rank 4, CUDA-graph replay) fail loudly on CPU tensors as well."""
    import headct_foundation_b200 as H
    from headct_foundation_b200 import _cabi
    from headct_foundation_b200.optim import FusedAdamW
    with pytest.raises(RuntimeError):
        H.AttentionClassifier(96, 2, num_heads=2)(torch.zeros(2, 9, 96))
    with pytest.raises(RuntimeError):
        H.RMSNorm(64)(torch.zeros(3, 64))
    with pytest.raises(RuntimeError):
        H.LoraLinear(64, 64, r=8)(torch.zeros(3, 64))
    lin = torch.nn.Linear(4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        H.GraphedForward(lin, torch.zeros(2, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        H.GraphedTrainStep(lin, FusedAdamW(lin.parameters()), torch.zeros(2, 4))
    L = _cabi.lib()
    assert L.hct_lora_shuffle(None, None, None, 2, 9, 2, 12, 0, None) == 1             # head dim not a multiple of 8
    assert L.hct_pool_attention_fwd(None, None, None, None, 2, 9, 2, 48, 9, 1.0, None) == 1   # more than 8 queries
    assert L.hct_colnorm_stats(None, None, 8, 6, 1e-6, 0.1, None, None, None, None, None) == 1  # dim % 4
    assert L.hct_lora_shuffle(None, None, None, 0, 9, 2, 48, 0, None) == 0             # empty batch is a no-op
    opt = FusedAdamW(lin.parameters(), lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05)
    lr, wd, bc1, bc2s, step = opt._hyper_values(opt.param_groups[0], 3)                # what the captured launch reads
    assert (lr, wd, step) == (1e-3, 0.05, 3.0) and abs(bc1 - (1 - 0.9 ** 3)) < 1e-12 and abs(bc2s - (1 - 0.95 ** 3) ** 0.5) < 1e-12
    # fp32-mode entry points validate their arguments without touching the device
    assert L.hct_split3_bf16(None, 6, 1, 0, 1, None, 4, 6, 0, 0, None) == 1                # cols % 4
    assert L.hct_attention_f32_fwd(None, None, None, 1, 8, 2, 40, None) == 3               # unsupported head dim
    assert L.hct_gelu_f32(None, None, 0, None) == 0                                        # empty is a no-op
    with pytest.raises(ValueError):
        H.set_precision("fp64")
    with H.precision("fp32"):
        assert H.get_precision() == "fp32"
        with pytest.raises(RuntimeError):
            H.AttentionBlock(96, 192, 2)(torch.zeros(2, 9, 96))                            # fp32 mode has no CPU fallback either
    assert H.get_precision() == "bf16"


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "headct_foundation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle's noise", ""), f"{f} mentions the oracle"


def test_constructor_error_conventions():
    import headct_foundation_b200 as H
    with pytest.raises(ValueError):
        H.AttentionBlock(768, 3072, 12, dropout_rate=1.5)
    with pytest.raises(ValueError):
        H.AttentionBlock(770, 3072, 12)
    with pytest.raises(ValueError):
        H.ViT(3, 96, 12, hidden_size=770)
    with pytest.raises(ValueError):
        H.PatchEmbeddingBlock(3, 8, 12, 768, 12)                     # patch > image
    with pytest.raises(ValueError):
        H.PatchEmbeddingBlock(3, 96, 12, 768, 12, pos_embed="bogus")
    with pytest.raises(ValueError):
        H.PatchEmbeddingBlock(3, 96, 12, 768, 12, patch_embed="bogus")
    with pytest.raises(AssertionError):
        H.PatchEmbeddingBlock(3, 100, 12, 768, 12)                   # not divisible
    with pytest.raises(AssertionError):
        H.build_sincos_position_embedding((8, 8, 8), 770, 3)
    with pytest.raises(NotImplementedError):
        H.build_sincos_position_embedding((8,), 768, 1)
