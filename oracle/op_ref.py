"""ctypes front-end of `dfa_oracle.c` (test infrastructure only — see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "dfa_oracle.c")
LIB = os.path.join(HERE, "_build", "libdfa_oracle.so")
_lib = None


def build_oracle(force=False):
    """gcc -O2 with FP contraction OFF: the only fused operation in the restatement is
    the explicit fmaf() that mirrors the reference's compiled FFMA."""
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared",
                               "-fPIC", "-o", LIB, SRC, "-lm"])
    return LIB


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_oracle())
        _lib.dfa_oracle_distinct_rows.restype = ctypes.c_longlong
    return _lib


def _np(x, dt):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=dt)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _dims(feat, shape, loc, w):
    bs, num_feat, C = feat.shape
    K, L = shape.shape[:2]
    A, P = loc.shape[1:3]
    G = w.shape[5]
    assert loc.shape == (bs, A, P, K, 2), loc.shape
    assert w.shape == (bs, A, P, K, L, G), w.shape
    return bs, num_feat, C, K, L, A, P, G


def forward(feat, spatial_shape, scale_start_index, loc, weights, fma_mode=1, side_channel=False):
    """Returns out[bs,A,C] float64 (and, with side_channel, valid[bs,A,P,K] uint8 and
    corner_rows[bs,A,P,K,L,4] int32)."""
    feat, loc, weights = _np(feat, np.float32), _np(loc, np.float32), _np(weights, np.float32)
    shape, start = _np(spatial_shape, np.int32), _np(scale_start_index, np.int32)
    bs, num_feat, C, K, L, A, P, G = _dims(feat, shape, loc, weights)
    out = np.empty((bs, A, C), np.float64)
    valid = np.zeros((bs, A, P, K), np.uint8) if side_channel else None
    rows = np.full((bs, A, P, K, L, 4), -1, np.int32) if side_channel else None
    rc = _load().dfa_oracle_forward(_ptr(feat), _ptr(shape), _ptr(start), _ptr(loc), _ptr(weights),
                                    _ptr(out), _ptr(valid), _ptr(rows), bs, num_feat, C, K, L, A, P,
                                    G, int(fma_mode))
    if rc:
        raise ValueError("dfa_oracle_forward rc=%d" % rc)
    return (out, valid, rows) if side_channel else out


def backward(feat, spatial_shape, scale_start_index, loc, weights, grad_out, fma_mode=1,
             need_feat=True):
    """Returns (grad_feat, grad_loc, grad_weights) in float64 (grad_feat None if not needed)."""
    feat, loc, weights = _np(feat, np.float32), _np(loc, np.float32), _np(weights, np.float32)
    go = _np(grad_out, np.float32)
    shape, start = _np(spatial_shape, np.int32), _np(scale_start_index, np.int32)
    bs, num_feat, C, K, L, A, P, G = _dims(feat, shape, loc, weights)
    assert go.shape == (bs, A, C)
    gf = np.empty(feat.shape, np.float64) if need_feat else None
    gl = np.empty(loc.shape, np.float64)
    gw = np.empty(weights.shape, np.float64)
    rc = _load().dfa_oracle_backward(_ptr(feat), _ptr(shape), _ptr(start), _ptr(loc), _ptr(weights),
                                     _ptr(go), _ptr(gf), _ptr(gl), _ptr(gw), bs, num_feat, C, K, L,
                                     A, P, G, int(fma_mode))
    if rc:
        raise ValueError("dfa_oracle_backward rc=%d" % rc)
    return gf, gl, gw


def distinct_rows(spatial_shape, scale_start_index, loc, num_feat, fma_mode=1):
    """U of SURVEY.md §8(d): distinct (b,row) pairs read by in-bounds corners of valid samples."""
    loc = _np(loc, np.float32)
    shape, start = _np(spatial_shape, np.int32), _np(scale_start_index, np.int32)
    bs, A, P, K, _ = loc.shape
    L = shape.shape[1]
    return int(_load().dfa_oracle_distinct_rows(_ptr(shape), _ptr(start), _ptr(loc), bs,
                                                int(num_feat), K, L, A, P, int(fma_mode)))
