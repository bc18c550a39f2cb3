"""BASELINE.json configs #4 and #5: the op at R101 1408x512 maps (bs=1, 8) and the sweep
anchors 900 -> 3600 x key points 13 -> 32 x fp32 / bf16 features (bs=1, R50 maps), forward and
backward, against the unmodified reference CUDA op (oracle/_ref, fp32 only) on the same inputs.
Cold L2 (rotating input sets), CUDA events inside a CUDA graph.  One JSON line per point.
Every point is also a PARITY point: forward output and the three gradients are compared with the
reference binary on the same inputs (max|x - ref| / max|ref| <= 1e-5; 2e-5 for the location gradient,
whose reference value comes from 1024-way float atomics) and the script exits non-zero on a miss.
bfloat16 points are checked against the reference op run on the bf16-rounded table.
    python tools/op_sweep.py [--quick]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import oracle  # noqa: E402
from simpb_b200 import cabi, synthetic  # noqa: E402

quick = "--quick" in sys.argv
peak, _ = bench.peaks()
ref = None
try:
    from oracle import build_ref
    if os.path.exists(build_ref.so_path()):
        ref = build_ref.load()
except Exception:
    ref = None


def rel(x, r):
    return float((x.double() - r.double()).abs().max() / r.double().abs().max().clamp_min(1e-30))


def parity(g):
    """Forward + three gradients of our op against the reference binary on one input set."""
    if ref is None or g["loc"].shape[0] * g["loc"].shape[1] * g["loc"].shape[2] > 349525:
        return None
    f32 = g["feat"].float()
    out = cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"])
    gf, gl, gw = cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"])
    rout = ref.deformable_aggregation_forward(f32, g["shape"], g["start"], g["loc"], g["w"])
    rgf, rgl, rgw = torch.zeros_like(f32), torch.zeros_like(g["loc"]), torch.zeros_like(g["w"])
    ref.deformable_aggregation_backward(f32, g["shape"], g["start"], g["loc"], g["w"], g["go"], rgf, rgl, rgw)
    e = {"out": rel(out, rout), "grad_feat": rel(gf, rgf), "grad_loc": rel(gl, rgl), "grad_w": rel(gw, rgw)}
    del rgf, gf
    ok = e["out"] <= 1e-5 and e["grad_feat"] <= 1e-5 and e["grad_w"] <= 1e-5 and e["grad_loc"] <= 2e-5
    return dict(rel_err=e, ok=ok)


FAILED = []


def point(name, levels, bs, A, P, dt):
    dtype = torch.float32 if dt == "f32" else torch.bfloat16
    esz = 4 if dt == "f32" else 2
    n_sets = 2 if (bs > 1 or levels is synthetic.R101_LEVELS) else 4
    host = [synthetic.rig_op_inputs(bs=bs, A=A, P=P, levels=levels, seed=s) for s in range(n_sets)]
    sets = [bench.to_device(d, dtype) for d in host]
    outs = [torch.empty(bs, A, 256, device="cuda") for _ in sets]
    sync = torch.cuda.synchronize
    fwd = [(lambda g=g, o=o: cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], out=o))
           for g, o in zip(sets, outs)]
    t_f = bench.time_graph(fwd, 60, 6, True, sync) / 60
    gf = torch.empty(sets[0]["feat"].shape, device="cuda", dtype=torch.float32)
    bwd = [(lambda g=g: cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"], gf,
                                      torch.empty_like(g["loc"]), torch.empty_like(g["w"]),
                                      flags=cabi.BWD_OVERWRITE_SMALL | cabi.BWD_ZERO_GRAD_FEAT)) for g in sets]
    t_b = bench.time_graph(bwd, 32, 4, True, sync) / 32     # a multiple of the set count: every launch is a graph replay
    u = oracle.distinct_rows(host[0]["spatial_shape"], host[0]["scale_start_index"],
                             host[0]["sampling_location"], host[0]["num_feat"])
    b_alg, _ = bench.algorithmic_bytes(host[0], esz, u)
    rec = {"point": name, "bs": bs, "anchors": A, "pts": P, "feat": dt, "num_feat": host[0]["num_feat"],
           "fwd_us": round(t_f * 1e3, 2), "bwd_us_incl_fill": round(t_b * 1e3, 2),
           "queries_per_s": round(bs * A / (t_f * 1e-3)), "alg_MB": round(b_alg / 1e6, 1),
           "fwd_GBps": round(b_alg / (t_f * 1e-3) / 1e9), "fwd_frac_of_measured_hbm": round(b_alg / (t_f * 1e-3) / 1e9 / peak, 3)}
    if ref is not None and dt == "f32" and bs * A * P <= 349525:     # int32 thread index of the reference
        rf = [(lambda g=g: ref.deformable_aggregation_forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"]))
              for g in sets]
        rec["reference_op_fwd_us"] = round(bench.time_graph(rf, 12, 2, False, sync) / 12 * 1e3, 1)
        g0 = sets[0]
        rg = [torch.zeros_like(g0["feat"]), torch.zeros_like(g0["loc"]), torch.zeros_like(g0["w"])]
        rb = [(lambda g=g: (rg[0].zero_(), rg[1].zero_(), rg[2].zero_(),
                            ref.deformable_aggregation_backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"],
                                                                g["go"], rg[0], rg[1], rg[2]))) for g in sets]
        rec["reference_op_bwd_us"] = round(bench.time_graph(rb, 6, 1, False, sync) / 6 * 1e3, 1)
        rec["fwd_speedup_vs_reference_op"] = round(rec["reference_op_fwd_us"] / rec["fwd_us"], 1)
        rec["bwd_speedup_vs_reference_op"] = round(rec["reference_op_bwd_us"] / rec["bwd_us_incl_fill"], 1)
    par = parity(sets[0])
    if par is not None:
        rec["parity_vs_reference_op"] = par
        if not par["ok"]:
            FAILED.append((name, bs, A, P, dt, par["rel_err"]))
    print(json.dumps(rec), flush=True)
    del sets, outs, gf
    torch.cuda.empty_cache()


R50, R101 = synthetic.R50_LEVELS, synthetic.R101_LEVELS
point("R50 bs1", R50, 1, 900, 13, "f32")
point("R50 bs8", R50, 8, 900, 13, "f32")
point("R50 bs8 train anchors", R50, 8, 1220, 13, "f32")
point("R101 bs1", R101, 1, 900, 13, "f32")
point("R101 bs8", R101, 8, 900, 13, "f32")
if not quick:
    for A in (900, 1800, 3600):
        for P in (13, 20, 32):
            for dt in ("f32", "bf16"):
                if (A, P, dt) == (900, 13, "f32"):
                    continue
                point("sweep", R50, 1, A, P, dt)
if FAILED:
    print("PARITY FAILED:", FAILED, file=sys.stderr)
    sys.exit(1)
