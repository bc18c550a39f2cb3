"""Oracle of multi-scale deformable attention (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

Two independent restatements of mmcv-full 1.7.1's `MultiScaleDeformableAttnFunction` (the reference
calls it at projects/mmdet3d_plugin/models/group_attn.py:229-233; mmcv is not vendored and cannot be
installed here, so there are no vectors from mmcv itself; both restatements are pinned to a fixture
from an independent third-party implementation of the same function, HuggingFace transformers'
MultiScaleDeformableAttention — tests/golden/msda_hf.npz, tests/test_msda.py):

* `forward` / `backward`: ctypes front-end of msda_oracle.c (mmcv's CUDA kernel, restated in C);
* `msda_grid_sample`: mmcv's own CPU formulation `multi_scale_deformable_attn_pytorch`
  (mmcv/ops/multi_scale_deform_attn.py) — per level a `grid_sample(bilinear, zeros,
  align_corners=False)` of the value map at `2*loc-1`, weighted sum over levels and points — written
  with this image's torch.  Differentiable, used for the gradient cross-check.
The two are compared with each other in tests/test_msda.py before either is trusted.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "msda_oracle.c")
LIB = os.path.join(HERE, "_build", "libmsda_oracle.so")
_lib = None


def build_oracle(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                               "-o", LIB, SRC, "-lm"])
    return LIB


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_oracle())
    return _lib


def _np(x, dt):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=dt)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _dims(value, shapes, loc):
    bs, S, M, D = value.shape
    Q, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    assert loc.shape == (bs, Q, M, L, P, 2) and shapes.shape == (L, 2)
    return bs, S, M, D, Q, L, P


def forward(value, spatial_shapes, level_start_index, loc, w, fma_mode=1):
    """value [bs,S,M,D], loc [bs,Q,M,L,P,2], w [bs,Q,M,L,P] → out [bs,Q,M*D] float64."""
    value, loc, w = _np(value, np.float32), _np(loc, np.float32), _np(w, np.float32)
    shapes, start = _np(spatial_shapes, np.int32), _np(level_start_index, np.int32)
    bs, S, M, D, Q, L, P = _dims(value, shapes, loc)
    out = np.empty((bs, Q, M * D), np.float64)
    _load().msda_oracle_forward(_ptr(value), _ptr(shapes), _ptr(start), _ptr(loc), _ptr(w), _ptr(out),
                                bs, S, M, D, Q, L, P, int(fma_mode))
    return out


def backward(value, spatial_shapes, level_start_index, loc, w, grad_out, fma_mode=1, need_value=True):
    value, loc, w = _np(value, np.float32), _np(loc, np.float32), _np(w, np.float32)
    go = _np(grad_out, np.float32)
    shapes, start = _np(spatial_shapes, np.int32), _np(level_start_index, np.int32)
    bs, S, M, D, Q, L, P = _dims(value, shapes, loc)
    gv = np.empty(value.shape, np.float64) if need_value else None
    gl = np.empty(loc.shape, np.float64)
    gw = np.empty(w.shape, np.float64)
    _load().msda_oracle_backward(_ptr(value), _ptr(shapes), _ptr(start), _ptr(loc), _ptr(w), _ptr(go),
                                 _ptr(gv), _ptr(gl), _ptr(gw), bs, S, M, D, Q, L, P, int(fma_mode))
    return gv, gl, gw


def msda_grid_sample(value, spatial_shapes, loc, w):
    """mmcv's `multi_scale_deformable_attn_pytorch`, restated.  value [bs,S,M,D] torch tensor."""
    bs, _, M, D = value.shape
    _, Q, _, L, P, _ = loc.shape
    sizes = [(int(h), int(ww)) for h, ww in spatial_shapes.tolist()]
    levels = value.split([h * ww for h, ww in sizes], dim=1)
    grids = 2 * loc - 1
    sampled = []
    for l, (h, ww) in enumerate(sizes):
        v = levels[l].flatten(2).transpose(1, 2).reshape(bs * M, D, h, ww)
        g = grids[:, :, :, l].transpose(1, 2).flatten(0, 1)               # [bs*M, Q, P, 2]
        sampled.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    att = w.transpose(1, 2).reshape(bs * M, 1, Q, L * P)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * att).sum(-1).view(bs, M * D, Q)
    return out.transpose(1, 2).contiguous()
