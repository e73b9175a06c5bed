"""CPU tests of the boundary: the C-ABI library builds/loads and exports every symbol include/evt.h declares,
argument validation works without a GPU, and the product path refuses to run without CUDA."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from edgevisiontransformer_b200 import _lib
    return _lib.load()


def test_header_symbols_exported(lib):
    from edgevisiontransformer_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "evt.h")).read()
    declared = set(re.findall(r"\b(evt_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"evt_model_spec", "evt_tensor_view"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/evt.h but not exported by libevt.so"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.evt_version() == 100


def test_struct_layout_matches_header():
    from edgevisiontransformer_b200 import _lib
    assert C.sizeof(_lib.ModelSpec) == 4 * 10 + 4 * 64 * 2 + 4 * 6        # ... embed_k, head_rows
    assert C.sizeof(_lib.TensorView) == 8 + 8 + 8 + 32


@pytest.mark.skipif(torch.cuda.is_available(), reason="exercises the no-GPU failure path")
def test_no_silent_cpu_fallback(lib):
    from edgevisiontransformer_b200 import B200ViTConfig, B200ViTForImageClassification, ops
    rc = lib.evt_device_check()
    assert rc != 0 and lib.evt_last_error()
    with pytest.raises(RuntimeError):
        ops.layernorm(torch.zeros(2, 8), torch.ones(8), torch.zeros(8), 1e-5)
    with pytest.raises(RuntimeError):
        B200ViTForImageClassification(B200ViTConfig(), {}, device="cpu")
    rc = lib.evt_gemm_bias_act(None, 8, None, 8, None, None, 0, 0, 0, None, 0, 8, 0, 0, 0, 1, 1, 1, 0, None)
    assert rc != 0
    from edgevisiontransformer_b200.modeling_swin import B200SwinForImageClassification
    with pytest.raises(RuntimeError):
        B200SwinForImageClassification({}, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], embed_dim=96, device="cpu")
    with pytest.raises(RuntimeError):
        ops.window_attention(torch.zeros(49, 288, dtype=torch.bfloat16), torch.zeros(1, 3, 64, 56), 1, 3)
    assert lib.evt_window_attention_fwd(None, 0, None, 0, None, 1, 1, 49, 3, 32, 0.0, None) != 0


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "edgevisiontransformer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
