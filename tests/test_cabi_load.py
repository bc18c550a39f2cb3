"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the header
declares, validates arguments without touching a GPU, and the product package never reaches
into the oracle."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_build_products_exist():
    from simpb_b200 import build
    lib, ext = build.build_all()
    assert os.path.exists(lib) and os.path.exists(ext)


def test_library_exports_every_declared_symbol():
    from simpb_b200 import cabi
    header = open(os.path.join(ROOT, "include", "dfa_b200.h")).read()
    declared = set(re.findall(r"\b(dfa_[a-z_]+)\s*\(", header))
    assert declared == set(cabi.SYMBOLS), declared ^ set(cabi.SYMBOLS)
    raw = ctypes.CDLL(cabi.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert cabi.lib.dfa_version() == 100


def test_argument_validation_needs_no_gpu():
    from simpb_b200 import cabi
    d = cabi.Dims(1, 6, 100, 256, 4, 9, 13, 8)
    rc = cabi.lib.dfa_forward(None, 0, None, None, None, None, None, ctypes.byref(d), None)
    assert rc == -1 and b"null" in cabi.lib.dfa_error_string(rc)
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    bad = cabi.Dims(1, 6, 100, 250, 4, 9, 13, 8)          # C % G != 0
    assert cabi.lib.dfa_forward(p, 0, p, p, p, p, p, ctypes.byref(bad), None) == -2
    zero = cabi.Dims(0, 6, 100, 256, 4, 9, 13, 8)          # empty batch
    assert cabi.lib.dfa_forward(p, 0, p, p, p, p, p, ctypes.byref(zero), None) == -2
    big = cabi.Dims(1, 6, 1 << 24, 256, 4, 9, 13, 8)       # 32-bit offset overflow
    assert cabi.lib.dfa_forward(p, 0, p, p, p, p, p, ctypes.byref(big), None) == -2
    assert cabi.lib.dfa_forward(p, 7, p, p, p, p, p, ctypes.byref(d), None) == -3
    assert cabi.lib.dfa_forward(p + 2, 0, p, p, p, p, p, ctypes.byref(d), None) == -4
    assert cabi.lib.dfa_forward_host_workspace_bytes(0, ctypes.byref(d)) > 4 * 100 * 256
    # host entry point and its statistics: same validation, before any CUDA call
    assert cabi.lib.dfa_forward_host(p, 0, p, p, p, p, p, ctypes.byref(d), None, 1 << 30, None) == -1
    assert cabi.lib.dfa_forward_host(p, 0, p, p, p, p, p, ctypes.byref(d), p, 16, None) == -2      # workspace too small
    assert cabi.lib.dfa_forward_host(p, 9, p, p, p, p, p, ctypes.byref(d), p, 1 << 30, None) == -3
    n = ctypes.c_int64(0)
    assert cabi.lib.dfa_forward_host_stats(None, 0, ctypes.byref(d), None, ctypes.byref(n), None, None) == -1
    assert cabi.lib.dfa_forward_host_stats(p, 0, ctypes.byref(bad), None, ctypes.byref(n), None, None) == -2


def test_argument_validation_of_the_other_entry_points():
    """Every entry point refuses null pointers and nonsense sizes with a DFA_ERR_* code before
    touching the device."""
    from simpb_b200 import cabi
    lib = cabi.lib
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    d = cabi.Dims(1, 6, 100, 256, 4, 9, 13, 8)
    assert lib.dfa_backward(p, 0, p, p, p, p, None, p, p, p, ctypes.byref(d), 0, None) == -1
    assert lib.dfa_backward(p, 0, p, p, p, p, p, None, p, p, ctypes.byref(cabi.Dims(1, 6, 100, 250, 4, 9, 13, 8)),
                            0, None) == -2                       # NULL grad_feat is allowed; dims are not
    assert lib.dfa_forward_fused(p, 0, p, p, None, p, 7, p, p, p, p, p, p, None, ctypes.byref(d), None) == -1
    assert lib.dfa_forward_fused(p, 0, p, p, p, p, 20, p, p, p, p, p, p, None, ctypes.byref(d), None) == -2
    odd_groups = cabi.Dims(1, 6, 100, 255, 4, 9, 13, 5)          # G not a power of two: caller falls back
    assert lib.dfa_forward_fused(p, 0, p, p, p, p, 7, p, p, p, p, p, p, None, ctypes.byref(odd_groups), None) == -5
    assert lib.dfa_debug_indices(p, p, None, p, p, ctypes.byref(d), None) == -1
    assert lib.dfa_flatten_maps(None, p, 4, 1, 6, 256, p, 0, None) == -1
    assert lib.dfa_flatten_maps(p, p, 0, 1, 6, 256, p, 0, None) == -2
    assert lib.dfa_keypoints_project(p, p, 7, None, p, p, None, p, 1, 9, 13, 6, None) == -1   # learnable pts need logits
    assert lib.dfa_keypoints_project(p, p, 14, p, p, p, None, p, 1, 9, 13, 6, None) == -2     # more fixed than total
    assert lib.dfa_keypoints_project_backward(p, p, 7, p, p, p, None, p, p, 1, 9, 13, 6, None) == -1
    assert lib.dfa_softmax_weights(None, None, 1.0, p, 1, 9, 6, 4, 13, 8, None) == -1
    assert lib.dfa_softmax_weights(p, None, 1.0, p, 1, 9, 6, 4, 13, 7, None) == -5            # 256 % G != 0
    assert lib.dfa_softmax_weights_split(p, None, None, 1.0, p, 1, 9, 6, 4, 13, 8, None) == -1
    assert lib.dfa_softmax_weights_backward(p, None, 1.0, p, None, 1, 9, 6, 4, 13, 8, None) == -1
    assert lib.dfa_softmax_weights_split_backward(p, p, None, 1.0, p, p, None, 1, 9, 6, 4, 13, 8, None) == -1
    assert lib.dfa_msda_forward(p, 0, p, p, p, None, p, 1, 100, 8, 32, 9, 4, 4, 1, None, None) == -1
    assert lib.dfa_msda_forward(p, 0, p, p, p, p, p, 1, 100, 8, 0, 9, 4, 4, 1, None, None) == -2
    assert lib.dfa_msda_forward(p, 0, p, p, p, p, p, 1, 100, 8, 32, 9, 4, 4, 3, None, None) == -1   # groups need a table
    assert lib.dfa_msda_forward(p, 9, p, p, p, p, p, 1, 100, 8, 32, 9, 4, 4, 1, None, None) == -3
    assert lib.dfa_msda_backward(p, 0, p, p, p, p, p, None, None, p, 1, 100, 8, 32, 9, 4, 4, 1, None, 1, None) == -1
    for code in (-1, -2, -3, -4, -5):
        assert lib.dfa_error_string(code).startswith(b"dfa:")


def test_cpu_tensors_are_rejected_loudly():
    from simpb_b200 import cabi, deformable_aggregation_function
    t = torch.zeros(1, 4, 8)
    with pytest.raises(cabi.DfaError):
        cabi.forward(t, torch.zeros(1, 1, 2, dtype=torch.int32), torch.zeros(1, 1, dtype=torch.int32),
                     torch.zeros(1, 1, 1, 1, 2), torch.zeros(1, 1, 1, 1, 1, 2))
    with pytest.raises(cabi.DfaError):
        deformable_aggregation_function(t, torch.zeros(1, 1, 2), torch.zeros(1, 1),
                                        torch.zeros(1, 1, 1, 1, 2), torch.zeros(1, 1, 1, 1, 1, 2))


def test_torch_extension_has_reference_entry_points():
    from simpb_b200.ops import deformable_aggregation_ext as ext
    assert callable(ext.deformable_aggregation_forward)
    assert callable(ext.deformable_aggregation_backward)
    with pytest.raises(RuntimeError):
        ext.deformable_aggregation_forward(*[torch.zeros(1)] * 5)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "simpb_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "oracle/" not in src and "dfa_oracle" not in src, f
