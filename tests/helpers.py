"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerances: 1e-5 relative (fp32), 1e-2 relative (bf16 features); "relative" is
# taken per tensor against max|ref|.
RTOL_F32 = 1e-5
RTOL_BF16 = 1e-2


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rel_err(x, ref):
    x = np.asarray(x.detach().cpu() if hasattr(x, "detach") else x, dtype=np.float64)
    ref = np.asarray(ref.detach().cpu() if hasattr(ref, "detach") else ref, dtype=np.float64)
    assert x.shape == ref.shape, (x.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    scale = np.abs(ref).max()
    if scale == 0:
        return float(np.abs(x).max())
    return float(np.abs(x - ref).max() / scale)


def assert_close(x, ref, rtol, what=""):
    e = rel_err(x, ref)
    assert e <= rtol, "%s: max|x-ref|/max|ref| = %.3e > %.1e" % (what, e, rtol)


def golden_op_inputs(g):
    """Flattened op inputs from an op_* golden file (maps → [bs, num_feat, C] by our own
    flattening restatement; the layout itself is pinned by flatten_small)."""
    from oracle import module_ref
    n_levels = g["sizes"].shape[0]
    maps = [torch.from_numpy(g["map%d" % l]) for l in range(n_levels)]
    col, shape, start = module_ref.flatten_feature_maps(maps)
    gmaps = [torch.from_numpy(g["grad_map%d" % l]) for l in range(n_levels)]
    gcol, _, _ = module_ref.flatten_feature_maps(gmaps)
    return col, shape, start, gcol


def seeded_state_dict(module, seed):
    """Deterministic parameters for a module whose checkpoint would be too large to commit: every
    floating-point entry of the state dict, in sorted key order, is drawn from one seeded generator
    (N(0,1) scaled by 1/sqrt(fan_in) for matrices, 1 + 0.1 N(0,1) for LayerNorm gains, 0.1 N(0,1) for
    other vectors); buffers such as fix_scale are left alone.  The golden generator and the tests both
    call this, so only the seed has to be stored."""
    gen = torch.Generator().manual_seed(int(seed))
    sd = module.state_dict()
    out = {}
    for k in sorted(sd):
        v = sd[k]
        if not v.is_floating_point() or k.endswith("fix_scale"):
            out[k] = v.clone()
        elif v.dim() >= 2:
            out[k] = torch.randn(v.shape, generator=gen) / float(v.shape[-1]) ** 0.5
        elif v.dim() == 1 and k.endswith("weight"):
            out[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=gen)
        else:
            out[k] = 0.1 * torch.randn(v.shape, generator=gen)
    return out
