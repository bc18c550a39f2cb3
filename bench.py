#!/usr/bin/env python
"""Benchmark of the deformable-feature-aggregation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload fwd|train] [--batch B] [--inputs rig|uniform] [--dtype f32|bf16]

One "step" = one pass of the op over one batch of synthetic SimPB-shaped input
(R50 704x256: 6 cameras x 4 FPN levels = 89,760 feature rows x 256 channels; 900 anchors x 13
key points x 8 groups).  Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for how every
number is defined.

* value       queries/s = bs*A*N / step time, inputs resident in HBM but cold in L2 (the step
              rotates over input sets whose total size exceeds the 126 MB L2), CUDA events on the
              launching stream, max over ranks.
* e2e         the same metric through the C ABI's host-buffer entry point (dfa_forward_host):
              pinned host inputs → device, kernel, output → host, every step.
* roofline    algorithmic bytes of SURVEY.md §8(d) / kernel time, against the measured HBM peak.
* cpu_baseline the reference's grid_sample CPU path (port under oracle/), bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "deformable_aggregation_forward_queries_per_sec"
UNIT = "queries/s"
L2_BYTES = 126e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="fwd", choices=["fwd", "train"])
    ap.add_argument("--batch", type=int, default=1, help="batch items per GPU per step")
    ap.add_argument("--anchors", type=int, default=900)
    ap.add_argument("--inputs", default="rig", choices=["rig", "uniform"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--no-graph", action="store_true", help="launch step by step (no CUDA graph)")
    ap.add_argument("--no-extras", action="store_true", help="skip variants / baselines")
    return ap.parse_args()


# ----------------------------------------------------------------------------- helpers
def measured_traffic(args):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, taken from the
    committed ncu capture of this exact workload (profiles/traffic.json); None if never captured."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    key = "%s:%s:%s:bs%d:A%d" % (args.workload, args.dtype, args.inputs, args.batch, args.anchors)
    if os.path.exists(p):
        t = json.load(open(p)).get(key)
        if t:
            return t["dram_bytes_read"] + t["dram_bytes_write"], t["kernel"]
    return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6
                          for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def make_inputs(args, seed, A=None, batch=None, levels=None):
    from simpb_b200 import synthetic
    maker = synthetic.rig_op_inputs if args.inputs == "rig" else synthetic.op_inputs_uniform
    kw = dict(bs=batch or args.batch, A=A or args.anchors, seed=seed)
    if levels is not None:
        kw["levels"] = levels
    return maker(**kw)


def to_device(d, dtype):
    return dict(feat=d["mc_ms_feat"].cuda().to(dtype).contiguous(),
                shape=d["spatial_shape"].int().cuda(), start=d["scale_start_index"].int().cuda(),
                loc=d["sampling_location"].cuda().contiguous(), w=d["weights"].cuda().contiguous(),
                go=d["grad_output"].cuda().contiguous())


def algorithmic_bytes(d, esz, distinct_rows):
    """forward, drop-in signature."""
    bs, A, P, K, _ = d["sampling_location"].shape
    L, G = d["weights"].shape[4:6]
    C = d["grad_output"].shape[-1]
    fixed = 8 * bs * A * P * K + 4 * bs * A * P * K * L * G + 4 * bs * A * C
    return fixed + distinct_rows * C * esz, fixed + bs * d["num_feat"] * C * esz


def time_graph(fn_list, steps, warmup, use_graph, sync_all, finalize=None):
    """Times `steps` launches cycling through fn_list (one callable per rotating input set).
    Returns total milliseconds measured by CUDA events on the current stream."""
    n = len(fn_list)
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
        for i in range(max(warmup, 3)):
            fn_list[i % n]()
        stream.synchronize()
        graph = None
        if use_graph:
            if finalize is not None:
                finalize()      # events recorded by the warm-up steps must not leak into the capture
                stream.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                for f in fn_list:
                    f()
                if finalize is not None:
                    finalize()  # side streams forked inside the capture (all-reduce) join it here
            graph.replay()
            stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record(stream)
        if graph is not None:
            for _ in range(steps // n):
                graph.replay()
            for i in range(steps % n):
                fn_list[i]()
        else:
            for i in range(steps):
                fn_list[i % n]()
        if finalize is not None:
            finalize()          # e.g. join the communication stream: its work belongs to the steps
        e1.record(stream)
        stream.synchronize()
        sync_all()
    return e0.elapsed_time(e1)


# ----------------------------------------------------------------------------- CPU baseline
def cpu_reference_setup(args):
    """The reference's CPU path = grid_sample per level + weighted fusion
    (/root/reference/projects/mmdet3d_plugin/models/blocks.py:148-156, :215-261), restated in
    oracle/module_ref.py.  One batch item of the bench workload per call."""
    from oracle import module_ref
    d = make_inputs(args, seed=0, batch=1)
    maps = module_ref.unflatten_feature_maps(d["mc_ms_feat"], d["spatial_shape"],
                                             d["scale_start_index"])
    uv = d["sampling_location"].permute(0, 3, 1, 2, 4).contiguous()        # [bs,K,A,P,2]
    w = d["weights"].permute(0, 1, 3, 4, 2, 5).contiguous()                # [bs,A,K,L,P,G]
    G = w.shape[-1]

    def step():
        with torch.no_grad():
            return module_ref.aggregate_grid_sample(maps, uv, w, G, apply_op_mask=False)
    return step, d["sampling_location"].shape[1]


def time_cpu(step, steps, warmup):
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return ts


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count())
    step, A = cpu_reference_setup(args)
    ts = time_cpu(step, args.steps, max(args.warmup, 1))
    ms = 1e3 * sum(ts) / len(ts)
    val = A / (ms / 1e3)
    sample = "1 batch item (%d anchors) of the %s workload per step, all host threads" % (A, args.inputs)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(args, note="reference CPU path (grid_sample), port under oracle/"),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(),
                             "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, note=None):
    c = {"workload": "deformable_aggregation %s, SimPB R50 704x256 shape: %d batch item(s)/GPU x "
                     "%d anchors x 13 key points x 6 cams x 4 levels x 8 groups, C=256"
                     % ("forward" if args.workload == "fwd" else "forward+backward",
                        args.batch, args.anchors),
         "inputs": ("S1 camera-rig geometry (~19% of samples valid)" if args.inputs == "rig"
                    else "S0 uniform locations in (-0.1,1.1) (~69% valid)"),
         "feature_dtype": args.dtype, "batch_per_gpu": args.batch, "anchors": args.anchors,
         "l2": "cold: steps rotate over input sets totalling more than the 126 MB L2"}
    if note:
        c["note"] = note
    return c


# ----------------------------------------------------------------------------- own arm
def run_own_arm(args):
    from simpb_b200 import cabi
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    dtype = torch.float32 if args.dtype == "f32" else torch.bfloat16
    esz = 4 if args.dtype == "f32" else 2
    host = [make_inputs(args, seed=1000 * rank + s) for s in range(1)]
    per_set = host[0]["mc_ms_feat"].numel() * esz + host[0]["weights"].numel() * 4
    n_sets = max(2, int(4 * L2_BYTES // per_set) + 1)
    n_sets = min(n_sets, 8)
    host += [make_inputs(args, seed=1000 * rank + s) for s in range(1, n_sets)]
    sets = [to_device(d, dtype) for d in host]
    outs = [torch.empty(args.batch, args.anchors, 256, device="cuda") for _ in sets]
    launches_per_step = 1
    finalize = None

    if args.workload == "fwd":
        fns = [(lambda g=g, o=o: cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], out=o))
               for g, o in zip(sets, outs)]
    else:
        # training step of the op: forward, backward (grad_feat memset inside), then the data-parallel
        # all-reduce (mean) of a gradient bucket the size of the three DFA layers' parameters
        # (3 x 247,495 fp32), enqueued on a side stream so it overlaps the next step's kernels
        from simpb_b200 import parallel
        gfs = [torch.empty_like(g["feat"], dtype=torch.float32) for g in sets[:2]]
        gls = [torch.empty_like(g["loc"]) for g in sets]
        gws = [torch.empty_like(g["w"]) for g in sets]
        params = [torch.nn.Parameter(torch.zeros(247495, device="cuda")) for _ in range(3)]
        for p in params:
            p.grad = torch.ones_like(p)
        bucket = None
        if dist is not None:
            bucket = parallel.GradBucket(params, comm_stream=torch.cuda.Stream())
            finalize = bucket.wait
        launches_per_step = 3          # forward, grad_feat memset, backward (+ NCCL's own kernels)

        fill_stream = torch.cuda.Stream()

        def mk(i, g, o):
            def f():
                # what DeformableAggregationFunction does: the zero-fill of the scatter target runs
                # on a side stream under the forward, the backward waits for it
                cur = torch.cuda.current_stream()
                fill_stream.wait_stream(cur)
                with torch.cuda.stream(fill_stream):
                    gfs[i % 2].zero_()
                    filled = torch.cuda.Event()
                    filled.record(fill_stream)
                cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], out=o)
                cur.wait_event(filled)
                cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"],
                              gfs[i % 2], gls[i], gws[i], flags=cabi.BWD_OVERWRITE_SMALL)
                if bucket is not None:
                    bucket.wait()                 # previous step's all-reduce must have landed
                    bucket.all_reduce_mean()
            return f
        fns = [mk(i, g, o) for i, (g, o) in enumerate(zip(sets, outs))]

    # The NCCL all-reduce is captured into the CUDA graph with the kernels (one launch per replay
    # instead of ~20 eager launches per step); DFA_BENCH_EAGER_DIST=1 keeps the multi-GPU training
    # step eager.
    use_graph = not args.no_graph and not (args.workload == "train" and dist is not None
                                           and os.environ.get("DFA_BENCH_EAGER_DIST") == "1")
    with ClockSampler(local) as clk:
        total_ms = time_graph(fns, args.steps, args.warmup, use_graph, sync_all, finalize)
    t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_step = total_ms / args.steps
    queries = args.batch * args.anchors * world
    value = queries / (ms_step / 1e3)

    line = None
    if rank == 0:
        import oracle
        peak, peak_src = peaks()
        u = [oracle.distinct_rows(d["spatial_shape"], d["scale_start_index"], d["sampling_location"],
                                  d["num_feat"]) for d in host]
        ab = [algorithmic_bytes(d, esz, ui) for d, ui in zip(host, u)]
        b_alg = sum(a for a, _ in ab) / len(ab)
        b_full = sum(f for _, f in ab) / len(ab)
        roof = None
        if args.workload == "fwd":
            ach = b_alg / (ms_step * 1e-3) / 1e9
            achf = b_full / (ms_step * 1e-3) / 1e9
            traffic, kname = measured_traffic(args)
            roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "kernel": kname or "dfa_fwd_rows_kernel",
                    "kernel_us": ms_step * 1e3,
                    "algorithmic_bytes": b_alg, "distinct_rows": sum(u) / len(u),
                    "peak_source": peak_src,
                    "whole_pyramid_variant": {"bytes": b_full, "achieved": achf, "frac": achf / peak}}
        else:
            # forward + backward of the op: the forward's bytes, plus for the backward
            # 2 x locations + 2 x weights + grad_out + the referenced rows once more + the zero fill of
            # grad_mc_ms_feat + the read-modify-write of the touched gradient rows
            d0 = host[0]
            bs, A, P, K, _ = d0["sampling_location"].shape
            L, G = d0["weights"].shape[4:6]
            C, U = 256, sum(u) / len(u)
            b_bwd = (2 * 8 * bs * A * P * K + 2 * 4 * bs * A * P * K * L * G + 4 * bs * A * C
                     + U * C * esz + bs * d0["num_feat"] * C * 4 + 2 * U * C * 4)
            ach = (b_alg + b_bwd) / (ms_step * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": None, "kernel": "dfa_fwd_rows_kernel + grad fill + dfa_bwd_merge_kernel",
                    "kernel_us": ms_step * 1e3, "algorithmic_bytes": b_alg + b_bwd,
                    "distinct_rows": U, "peak_source": peak_src,
                    "note": "whole training step of the op (all-reduce overlapped when n_gpus > 1)"}
        line = {"metric": METRIC if args.workload == "fwd" else "deformable_aggregation_fwd_bwd_queries_per_sec",
                "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": workload_config(args), "clocks": clk.summary(),
                "gpu_launches": launches_per_step * args.steps, "cuda_graph": bool(use_graph),
                "roofline": roof}

    # ---- end to end through the host-buffer C ABI entry point (every rank, max over ranks)
    d0 = host[0]
    dims = cabi.Dims(args.batch, 6, d0["num_feat"], 256, 4, args.anchors, 13, 8)
    hf = cabi.HostForward(dims, dtype)
    pin = lambda x: x.contiguous().pin_memory()  # noqa: E731
    h = [dict(feat=pin(d["mc_ms_feat"].to(dtype)), shape=pin(d["spatial_shape"].int()),
              start=pin(d["scale_start_index"].int()), loc=pin(d["sampling_location"]),
              w=pin(d["weights"])) for d in host[:2]]
    h_out = torch.empty(args.batch, args.anchors, 256).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    for i in range(3):
        hf(h[i % 2]["feat"], h[i % 2]["shape"], h[i % 2]["start"], h[i % 2]["loc"], h[i % 2]["w"], h_out)
    sync_all()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        x = h[i % 2]
        hf(x["feat"], x["shape"], x["start"], x["loc"], x["w"], h_out)   # ends with a stream sync
    e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    if rank == 0:
        x = h[0]
        h2d = sum(x[k].numel() * x[k].element_size() for k in ("feat", "shape", "start", "loc", "w"))
        line["e2e"] = {"value": queries / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
                       "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": h_out.numel() * 4,
                       "api": "dfa_forward_host (C ABI, pinned host buffers)", "steps": e2e_steps}

    # ---- baselines and variants (rank 0, N=1 only)
    if rank == 0 and world == 1:
        torch.set_num_threads(os.cpu_count())
        step, A = cpu_reference_setup(args)
        ts = time_cpu(step, 8, 1)
        cpu_ms = 1e3 * statistics.median(ts)
        line["cpu_baseline"] = {
            "value": A / (cpu_ms / 1e3), "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": "port", "ms_per_forward": cpu_ms,
            "sample": "1 batch item (%d anchors), reference grid_sample path, median of 8 after 1 "
                      "warm-up" % A}
        if not args.no_extras:
            line["variants"] = extras(args, cabi, sets, host, peak)
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def extras(args, cabi, sets, host, peak):
    """Secondary measurements reported beside the headline (same timing method)."""
    out = {}
    nosync = torch.cuda.synchronize

    def t_of(fns, steps=100):
        return time_graph(fns, steps, 10, True, nosync) / steps

    fwd = [(lambda g=g: cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"])) for g in sets]
    try:
        # warm L2: one input set over and over
        ms = t_of(fwd[:1])
        out["fwd_warm_l2_us"] = ms * 1e3
        # backward (cold), grad_feat zero-fill included
        gf = torch.empty_like(sets[0]["feat"], dtype=torch.float32)
        bwd = [(lambda g=g: cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"], gf,
                                          torch.empty_like(g["loc"]), torch.empty_like(g["w"]),
                                          flags=cabi.BWD_OVERWRITE_SMALL | cabi.BWD_ZERO_GRAD_FEAT))
               for g in sets]
        out["bwd_cold_us"] = t_of(bwd, 40) * 1e3
        bwd_nz = [(lambda g=g: cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"], gf,
                                             torch.empty_like(g["loc"]), torch.empty_like(g["w"]),
                                             flags=cabi.BWD_OVERWRITE_SMALL)) for g in sets]
        out["bwd_cold_no_memset_us"] = t_of(bwd_nz, 40) * 1e3
    except Exception as e:  # pragma: no cover
        out["error"] = repr(e)
    # the unmodified reference CUDA op on this GPU (oracle/_ref, built by oracle/build_ref.py)
    try:
        from oracle import build_ref
        if os.path.exists(build_ref.so_path()) and args.dtype == "f32":
            ref = build_ref.load()
            rf = [(lambda g=g: ref.deformable_aggregation_forward(g["feat"], g["shape"], g["start"],
                                                                  g["loc"], g["w"])) for g in sets]
            ms = time_graph(rf, 40, 5, False, nosync) / 40
            out["reference_cuda_op_fwd_cold_us"] = ms * 1e3
            g = sets[0]
            gfr, glr, gwr = torch.zeros_like(g["feat"]), torch.zeros_like(g["loc"]), torch.zeros_like(g["w"])
            rb = [(lambda g=g: (gfr.zero_(), glr.zero_(), gwr.zero_(),
                                ref.deformable_aggregation_backward(g["feat"], g["shape"], g["start"],
                                                                    g["loc"], g["w"], g["go"], gfr, glr, gwr)))
                  for g in sets]
            out["reference_cuda_op_bwd_cold_us"] = time_graph(rb, 20, 3, False, nosync) / 20 * 1e3
    except Exception as e:  # pragma: no cover
        out["reference_cuda_op_error"] = repr(e)
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
