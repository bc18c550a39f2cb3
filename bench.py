#!/usr/bin/env python
"""Benchmark of the deformable-feature-aggregation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload fwd|train] [--batch B] [--anchors A] [--inputs rig|uniform]
                    [--dtype f32|bf16] [--reps R] [--no-extras]

One "step" = one pass of the op over one batch of synthetic SimPB-shaped input
(R50 704x256: 6 cameras x 4 FPN levels = 89,760 feature rows x 256 channels; 900 anchors x 13
key points x 8 groups).  Prints ONE JSON line (rank 0).  See DESIGN.md §5 for how every number is
defined.

* value       queries/s = bs*A*N / step time, inputs resident in HBM but cold in L2 (the step
              rotates over input sets whose total size exceeds the 126 MB L2), CUDA events on the
              launching stream.  The timed region of exactly K steps (barrier + synchronize on both
              sides, max over ranks) is repeated `--reps` times and the MEDIAN region is reported, so a
              20-launch region on 8 GPUs is not timer noise (`timing` holds min / median / max).
* e2e         the same metric through the C ABI's host-buffer entry point (dfa_forward_host):
              pinned host inputs -> device, kernel, output -> host, every step; per-rank values and
              achieved host-to-device GB/s beside it, plus the same with a bfloat16 table on the host.
* roofline    algorithmic bytes of SURVEY.md §8(d) / kernel time, against the measured HBM peak; also
              holds (keys the driver keeps) the bs=8 and R101 forward fractions, the unmodified
              reference CUDA op's times on the same GPU, and the `train` record: forward + backward at 8
              items per GPU x 1,220 anchors with the NCCL all-reduce of the DFA gradient bucket.
* cpu_baseline the reference's grid_sample CPU path on the host cores, bounded sample: the reference's
              own Python (staged under oracle/_ref/py by build()) when present, else the port.
"""
import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "deformable_aggregation_forward_queries_per_sec"
UNIT = "queries/s"
L2_BYTES = 126e6
TRAIN_ANCHORS = 1220          # 900 + up to 320 denoising anchors (models/simpb_head.py:371-379)
BUCKET_FLOATS = 3 * 247495    # parameters of the three DFA layers of a frame


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
    ap.add_argument("--reps", type=int, default=31, help="repetitions of the timed region (median reported)")
    ap.add_argument("--no-graph", action="store_true", help="launch step by step (no CUDA graph)")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary shapes / baselines")
    return ap.parse_args()


# ----------------------------------------------------------------------------- helpers
_SYN = None


def synthetic():
    """simpb_b200/synthetic.py loaded BY PATH: the input generator is plain torch, and importing the
    package would dlopen libdfa_b200.so — which the reference arm must not do."""
    global _SYN
    if _SYN is None:
        spec = importlib.util.spec_from_file_location("dfa_bench_synthetic",
                                                      os.path.join(ROOT, "simpb_b200", "synthetic.py"))
        _SYN = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_SYN)
    return _SYN


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


def make_inputs(args, seed, A=None, batch=None, levels=None, feat=True):
    syn = synthetic()
    maker = syn.rig_op_inputs if args.inputs == "rig" else syn.op_inputs_uniform
    kw = dict(bs=batch or args.batch, A=A or args.anchors, seed=seed, feat=feat)
    if levels is not None:
        kw["levels"] = levels
    return maker(**kw)


def to_device(d, dtype):
    if d["mc_ms_feat"] is None:   # secondary shapes: the table's values do not matter for the clock
        feat = torch.randn(d["sampling_location"].shape[0], d["num_feat"], 256, device="cuda").to(dtype)
    else:
        feat = d["mc_ms_feat"].cuda().to(dtype).contiguous()
    return dict(feat=feat, shape=d["spatial_shape"].int().cuda(), start=d["scale_start_index"].int().cuda(),
                loc=d["sampling_location"].cuda().contiguous(), w=d["weights"].cuda().contiguous(),
                go=d["grad_output"].cuda().contiguous())


def algorithmic_bytes(d, esz, distinct_rows):
    """forward, drop-in signature."""
    bs, A, P, K, _ = d["sampling_location"].shape
    L, G = d["weights"].shape[4:6]
    C = d["grad_output"].shape[-1]
    fixed = 8 * bs * A * P * K + 4 * bs * A * P * K * L * G + 4 * bs * A * C
    return fixed + distinct_rows * C * esz, fixed + bs * d["num_feat"] * C * esz


def backward_bytes(d, esz, U):
    """backward = 2 x locations + 2 x weights + grad_out + the referenced rows once more
    + the zero fill of grad_mc_ms_feat + the read-modify-write of the touched gradient rows."""
    bs, A, P, K, _ = d["sampling_location"].shape
    L, G = d["weights"].shape[4:6]
    C = 256
    return (2 * 8 * bs * A * P * K + 2 * 4 * bs * A * P * K * L * G + 4 * bs * A * C
            + U * C * esz + bs * d["num_feat"] * C * 4 + 2 * U * C * 4)


def time_regions(fn_list, steps, warmup, use_graph, sync_all, finalize=None, reps=1):
    """Times `reps` regions of exactly `steps` launches cycling through fn_list (one callable per
    rotating input set).  Returns the list of region times in milliseconds (CUDA events on the
    launching stream; `sync_all` = barrier + synchronize on both sides of every region)."""
    n = len(fn_list)
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    out = []
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
        for _ in range(max(reps, 1)):
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
            out.append(e0.elapsed_time(e1))
    return out


def time_graph(fn_list, steps, warmup, use_graph, sync_all, finalize=None):
    """One region (tools/ use this): total milliseconds of `steps` launches."""
    return time_regions(fn_list, steps, warmup, use_graph, sync_all, finalize, 1)[0]


# ----------------------------------------------------------------------------- CPU baseline
def cpu_reference_setup(args):
    """The reference's CPU path = grid_sample per level + weighted fusion
    (/root/reference/projects/mmdet3d_plugin/models/blocks.py:148-156, :215-261) on one batch item of
    the bench workload.  With the reference's own files staged (oracle/_ref/py, or /root/reference in
    the build container) the step calls the REFERENCE'S feature_sampling + multi_view_level_fusion
    (kind "reference"); otherwise the restatement in oracle/module_ref.py (kind "port")."""
    from oracle import module_ref, ref_import
    d = make_inputs(args, seed=0, batch=1)
    maps = module_ref.unflatten_feature_maps(d["mc_ms_feat"], d["spatial_shape"],
                                             d["scale_start_index"])
    w = d["weights"].permute(0, 1, 3, 4, 2, 5).contiguous()                # [bs,A,K,L,P,G]
    A, P, G = w.shape[1], w.shape[4], w.shape[-1]
    root = ref_import.default_root() if args.inputs == "rig" else None
    if root is not None:
        blocks, _ = ref_import.import_reference(root)
        DFA = blocks.DeformableFeatureAggregation
        me = types.SimpleNamespace(num_groups=G, group_dims=256 // G, num_pts=P, embed_dims=256)
        kp, proj, wh = d["key_points"], d["projection_mat"], d["image_wh"]

        def step():
            with torch.no_grad():
                f = DFA.feature_sampling(maps, kp, proj, wh)                # blocks.py:215-245
                f = DFA.multi_view_level_fusion(me, f, w)                   # blocks.py:247-261
                return f.sum(dim=2)                                         # blocks.py:156
        return step, A, "reference"
    uv = d["sampling_location"].permute(0, 3, 1, 2, 4).contiguous()        # [bs,K,A,P,2]

    def step():
        with torch.no_grad():
            return module_ref.aggregate_grid_sample(maps, uv, w, G, apply_op_mask=False)
    return step, A, "port"


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
    step, A, kind = cpu_reference_setup(args)
    ts = time_cpu(step, args.steps, max(args.warmup, 1))
    ms = 1e3 * sum(ts) / len(ts)
    val = A / (ms / 1e3)
    sample = ("op level (grid_sample x 4 levels + weighted fusion for given key points and weights; no "
              "weights_fc / key-point FCs): 1 batch item (%d anchors) of the %s workload per step, all host "
              "threads; %s" % (A, args.inputs,
                               "the reference's own feature_sampling + multi_view_level_fusion "
                               "(models/blocks.py:215-261)" if kind == "reference"
                               else "port of models/blocks.py:215-261 under oracle/"))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(),
                             "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args):
    return {"workload": "deformable_aggregation %s, SimPB R50 704x256 shape: %d batch item(s)/GPU x "
                        "%d anchors x 13 key points x 6 cams x 4 levels x 8 groups, C=256"
                        % ("forward" if args.workload == "fwd" else "forward+backward",
                           args.batch, args.anchors),
            "inputs": ("S1 camera-rig geometry (~19% of samples valid)" if args.inputs == "rig"
                       else "S0 uniform locations in (-0.1,1.1) (~69% valid)"),
            "feature_dtype": args.dtype, "batch_per_gpu": args.batch, "anchors": args.anchors,
            "l2": "cold: steps rotate over input sets totalling more than the 126 MB L2"}


# ----------------------------------------------------------------------------- own arm
def train_step_fns(cabi, sets, outs, bucket):
    """The training step of the op: forward, backward (the two small gradients written in full), then —
    with a bucket — the data-parallel all-reduce (mean) of a gradient bucket the size of the three DFA
    layers' parameters, enqueued on a side stream so that it overlaps the next step's kernels.
    grad_mc_ms_feat has to start from zero (different anchors meet on one pixel): that fill does not
    depend on the forward, so it runs on a forked stream NEXT TO the forward kernel — a DRAM-write stream
    beside a gather that leaves two thirds of the DRAM bandwidth idle — and joins before the backward
    (DFA_BENCH_SERIAL_FILL=1: the library's own fill in front of the backward kernel, as in round 1)."""
    gfs = [torch.empty_like(g["feat"], dtype=torch.float32) for g in sets[:1]]
    gls = [torch.empty_like(g["loc"]) for g in sets]
    gws = [torch.empty_like(g["w"]) for g in sets]
    serial = os.environ.get("DFA_BENCH_SERIAL_FILL", "0") == "1"
    side = torch.cuda.Stream()

    def mk(i, g, o):
        def f():
            if serial:
                cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], out=o)
                cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"],
                              gfs[0], gls[i], gws[i], flags=cabi.BWD_OVERWRITE_SMALL | cabi.BWD_ZERO_GRAD_FEAT)
            else:
                cur = torch.cuda.current_stream()
                side.wait_stream(cur)             # after the previous step's backward (it wrote gfs[0])
                with torch.cuda.stream(side):
                    gfs[0].zero_()
                cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], out=o)
                cur.wait_stream(side)
                cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"],
                              gfs[0], gls[i], gws[i], flags=cabi.BWD_OVERWRITE_SMALL)
            if bucket is not None:
                bucket.wait()                 # the previous step's all-reduce must have landed
                bucket.all_reduce_mean()
        return f
    return [mk(i, g, o) for i, (g, o) in enumerate(zip(sets, outs))]


def make_bucket(dist):
    from simpb_b200 import parallel
    params = [torch.nn.Parameter(torch.zeros(BUCKET_FLOATS // 3, device="cuda")) for _ in range(3)]
    for p in params:
        p.grad = torch.ones_like(p)
    # gradients live in the bucket (as in DDP): no pack / unpack copies around the collective
    return (parallel.GradBucket(params, comm_stream=torch.cuda.Stream(), alias_grads=True)
            if dist is not None else None)


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

    def max_over_ranks(values):
        t = torch.tensor(values, device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def gather_ranks(value):
        t = torch.tensor([value], device="cuda", dtype=torch.float64)
        if dist is None:
            return [float(value)]
        ts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(ts, t)
        return [float(x.item()) for x in ts]

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
        bucket = make_bucket(dist)
        if bucket is not None:
            finalize = bucket.wait
        launches_per_step = 3          # forward, grad_feat memset, backward (+ NCCL's own kernels)
        fns = train_step_fns(cabi, sets, outs, bucket)

    # The NCCL all-reduce is captured into the CUDA graph with the kernels (one launch per replay
    # instead of ~20 eager launches per step); DFA_BENCH_EAGER_DIST=1 keeps the multi-GPU training
    # step eager.
    use_graph = not args.no_graph and not (args.workload == "train" and dist is not None
                                           and os.environ.get("DFA_BENCH_EAGER_DIST") == "1")
    with ClockSampler(local) as clk:
        regions = max_over_ranks(time_regions(fns, args.steps, args.warmup, use_graph, sync_all, finalize,
                                              args.reps))
    total_ms = statistics.median(regions)
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
            U = sum(u) / len(u)
            b_bwd = backward_bytes(host[0], esz, U)
            ach = (b_alg + b_bwd) / (ms_step * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": None, "kernel": "forward + grad fill + dfa_bwd_merge_kernel",
                    "kernel_us": ms_step * 1e3, "algorithmic_bytes": b_alg + b_bwd,
                    "distinct_rows": U, "peak_source": peak_src,
                    "note": "whole training step of the op (all-reduce overlapped when n_gpus > 1)"}
        line = {"metric": METRIC if args.workload == "fwd" else "deformable_aggregation_fwd_bwd_queries_per_sec",
                "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": workload_config(args), "clocks": clk.summary(),
                "gpu_launches": launches_per_step * args.steps * max(args.reps, 1), "cuda_graph": bool(use_graph),
                "timing": {"regions": len(regions), "steps_per_region": args.steps,
                           "region_ms_min": min(regions), "region_ms_median": total_ms,
                           "region_ms_max": max(regions), "reported": "median region / steps, max over ranks"},
                "roofline": roof}

    # ---- end to end through the host-buffer C ABI entry point (every rank; the slowest rank counts)
    def e2e_leg(dt, pull=True):
        d0 = host[0]
        dims = cabi.Dims(args.batch, 6, d0["num_feat"], 256, 4, args.anchors, 13, 8)
        os.environ["DFA_HOST_PULL"] = "1" if pull else "0"
        cabi.reload_knobs()
        hf = cabi.HostForward(dims, dt)
        pin = lambda x: x.contiguous().pin_memory()  # noqa: E731
        h = [dict(feat=pin(d["mc_ms_feat"].to(dt)), shape=pin(d["spatial_shape"].int()),
                  start=pin(d["scale_start_index"].int()), loc=pin(d["sampling_location"]),
                  w=pin(d["weights"])) for d in host[:2]]
        h_out = torch.empty(args.batch, args.anchors, 256).pin_memory()
        n = max(3, min(args.steps, 20))
        for i in range(3):
            hf(h[i % 2]["feat"], h[i % 2]["shape"], h[i % 2]["start"], h[i % 2]["loc"], h[i % 2]["w"], h_out)
        sync_all()
        t0 = time.perf_counter()
        for i in range(n):
            x = h[i % 2]
            hf(x["feat"], x["shape"], x["start"], x["loc"], x["w"], h_out)   # ends with a stream sync
        ms = 1e3 * (time.perf_counter() - t0) / n
        h2d, rows, wbytes = hf.stats()      # what the last call moved host -> device (counted on the device)
        whole = sum(h[0][k].numel() * h[0][k].element_size() for k in ("feat", "shape", "start", "loc", "w"))
        per_rank = gather_ranks(ms)
        del hf, h
        os.environ.pop("DFA_HOST_PULL", None)
        cabi.reload_knobs()
        return {"ms_per_step": max(per_rank), "per_rank_ms": per_rank, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": h_out.numel() * 4, "steps": n,
                "host_operand_bytes": whole, "feature_rows_moved": rows, "weight_bytes_moved": wbytes,
                "h2d_GBps_per_rank": [h2d / (m * 1e-3) / 1e9 for m in per_rank],
                "h2d_GBps_all_ranks": sum(h2d / (m * 1e-3) / 1e9 for m in per_rank)}

    e = e2e_leg(dtype)
    e_copy = e2e_leg(dtype, pull=False)
    e_bf16 = e2e_leg(torch.bfloat16) if dtype != torch.bfloat16 else None
    if rank == 0:
        line["e2e"] = {"value": queries / (e["ms_per_step"] / 1e3), "unit": UNIT,
                       "api": "dfa_forward_host (C ABI, pinned host buffers; pull mode: the device reads the "
                              "feature rows and weight lines the forward references straight from the pinned "
                              "host buffers, everything else is copied whole; include/dfa_b200.h)", **e,
                       "limiter": "the host link: one PCIe Gen5 x16 per GPU; with several ranks also the host's "
                                  "memory / root-complex bandwidth (h2d_GBps_all_ranks)",
                       "cpu_affinity_cores": len(os.sched_getaffinity(0)),
                       "whole_copy": {"value": queries / (e_copy["ms_per_step"] / 1e3), "unit": UNIT, **e_copy,
                                      "note": "same call with DFA_HOST_PULL=0: all five inputs copied whole "
                                              "(the only mode of round 1)"}}
        if e_bf16 is not None:
            line["e2e"]["bf16_table"] = {"value": queries / (e_bf16["ms_per_step"] / 1e3), "unit": UNIT, **e_bf16,
                                         "note": "same call with a bfloat16 feature table on the host: "
                                                 "half the feature bytes over the link"}

    # ---- the training configuration on every rank (BASELINE.json config #3)
    if not args.no_extras and args.workload == "fwd":
        tr = train_record(args, cabi, dist, world, rank, sync_all, max_over_ranks)
        if rank == 0:
            line["roofline"]["train"] = tr

    # ---- baselines and secondary shapes (rank 0, N=1 only)
    if rank == 0 and world == 1:
        torch.set_num_threads(os.cpu_count())
        step, A, kind = cpu_reference_setup(args)
        ts = time_cpu(step, 8, 1)
        cpu_ms = 1e3 * statistics.median(ts)
        line["cpu_baseline"] = {
            "value": A / (cpu_ms / 1e3), "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": kind, "ms_per_forward": cpu_ms,
            "sample": "op level (grid_sample + fusion for given key points and weights): 1 batch item "
                      "(%d anchors), median of 8 after 1 warm-up; %s" % (
                          A, "the reference's own feature_sampling + multi_view_level_fusion"
                          if kind == "reference" else "port under oracle/")}
        if not args.no_extras:
            del sets, outs
            torch.cuda.empty_cache()
            extras(args, cabi, host, peak, line["roofline"])
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def train_record(args, cabi, dist, world, rank, sync_all, max_over_ranks):
    """BASELINE.json config #3 on the driver's clock: forward + backward at 8 batch items per GPU x
    1,220 anchors (900 + 320 denoising, models/simpb_head.py:371-379), followed by the NCCL
    all-reduce (mean) of the 3 x 247,495-float DFA gradient bucket (apis/mmdet_train.py:97-102).
    Reports the step with and without the collective (their difference is what data parallelism costs:
    the weak-scaling loss) and the all-reduce alone."""
    bs, A = 8, TRAIN_ANCHORS
    a2 = argparse.Namespace(**vars(args))
    a2.batch, a2.anchors, a2.workload = bs, A, "train"
    host = [make_inputs(a2, seed=7000 + 100 * rank + s, feat=False) for s in range(2)]
    sets = [to_device(d, torch.float32) for d in host]
    outs = [torch.empty(bs, A, 256, device="cuda") for _ in sets]
    steps, reps = 8, 9
    local_fns = train_step_fns(cabi, sets, outs, None)
    t_local = statistics.median(max_over_ranks(time_regions(local_fns, steps, 4, True, sync_all, None, reps)))
    rec = {"workload": "forward + backward, %d items/GPU x %d anchors, fp32, rig inputs, cold L2" % (bs, A),
           "n_gpus": world, "ms_per_step_no_allreduce": t_local / steps,
           "timing": "median of %d regions of %d steps, max over ranks" % (reps, steps)}
    if dist is not None:
        bucket = make_bucket(dist)
        # four steps per CUDA graph: the all-reduce of a step runs under the next step's kernels, only the
        # last one of a replay is exposed (a captured graph must join its side stream before it ends)
        fns = train_step_fns(cabi, sets, outs, bucket) * 2
        t = statistics.median(max_over_ranks(time_regions(fns, steps, 4, True, sync_all, bucket.wait, reps)))
        rec["ms_per_step"] = t / steps

        def ar():
            bucket.all_reduce_mean()
            bucket.wait()
        t_ar = statistics.median(max_over_ranks(time_regions([ar], 20, 5, True, sync_all, None, reps)))
        rec["allreduce_alone_us"] = t_ar / 20 * 1e3
        rec["allreduce_bytes"] = 4 * BUCKET_FLOATS
        rec["allreduce_cost_in_step_us"] = (t - t_local) / steps * 1e3
        rec["weak_scaling_efficiency_vs_no_collective"] = t_local / t
        rec["limiter"] = ("all-reduce latency: a 3 MB message is in NCCL's latency-bound regime; it runs on a "
                          "side stream under the next step's forward, what is left shows in "
                          "allreduce_cost_in_step_us")
    else:
        rec["ms_per_step"] = t_local / steps
    rec["queries_per_s"] = bs * A * world / (rec["ms_per_step"] * 1e-3)
    if rank == 0:
        import oracle
        U = oracle.distinct_rows(host[0]["spatial_shape"], host[0]["scale_start_index"],
                                 host[0]["sampling_location"], host[0]["num_feat"])
        b = algorithmic_bytes(host[0], 4, U)[0] + backward_bytes(host[0], 4, U)
        rec["algorithmic_bytes_per_step_per_gpu"] = b
        rec["frac_of_measured_hbm"] = b / (rec["ms_per_step"] * 1e-3) / 1e9 / peaks()[0]
    del sets, outs
    torch.cuda.empty_cache()
    return rec


def extras(args, cabi, host, peak, roof):
    """Secondary measurements written INTO the roofline object (the driver keeps that key): the
    forward at bs=8 and at R101 maps (BASELINE.json config #4), warm-L2 and backward times of the
    headline shape, and the unmodified reference CUDA op on the same GPU and inputs."""
    import oracle
    nosync = torch.cuda.synchronize
    syn = synthetic()

    def fwd_point(levels, bs, A, n_sets=2, steps=40):
        a2 = argparse.Namespace(**vars(args))
        hs = [make_inputs(a2, seed=9000 + s, A=A, batch=bs, levels=levels, feat=False) for s in range(n_sets)]
        ss = [to_device(d, torch.float32) for d in hs]
        os_ = [torch.empty(bs, A, 256, device="cuda") for _ in ss]
        fns = [(lambda g=g, o=o: cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], out=o))
               for g, o in zip(ss, os_)]
        ms = statistics.median(time_regions(fns, steps, 6, True, nosync, None, 7)) / steps
        U = oracle.distinct_rows(hs[0]["spatial_shape"], hs[0]["scale_start_index"], hs[0]["sampling_location"],
                                 hs[0]["num_feat"])
        b_alg = algorithmic_bytes(hs[0], 4, U)[0]
        rec = {"kernel_us": ms * 1e3, "algorithmic_bytes": b_alg, "achieved_GBps": b_alg / (ms * 1e-3) / 1e9,
               "frac": b_alg / (ms * 1e-3) / 1e9 / peak, "queries_per_s": bs * A / (ms * 1e-3)}
        return rec, ss

    other = {}
    try:
        other["r50_bs8"], ss8 = fwd_point(syn.R50_LEVELS, 8, 900)
        # the unmodified reference CUDA op (oracle/_ref, built by oracle/build_ref.py) at bs=8
        ref = None
        try:
            from oracle import build_ref
            if os.path.exists(build_ref.so_path()):
                ref = build_ref.load()
        except Exception as e:  # pragma: no cover
            other["reference_cuda_op_error"] = repr(e)
        if ref is not None:
            rf = [(lambda g=g: ref.deformable_aggregation_forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"]))
                  for g in ss8]
            other["r50_bs8"]["reference_cuda_op_us"] = time_graph(rf, 6, 2, False, nosync) / 6 * 1e3
        del ss8
        torch.cuda.empty_cache()
        other["r50_bs8_train_anchors"], s_ = fwd_point(syn.R50_LEVELS, 8, TRAIN_ANCHORS)
        del s_
        torch.cuda.empty_cache()
        other["r101_bs1"], s_ = fwd_point(syn.R101_LEVELS, 1, 900, n_sets=2, steps=60)
        del s_
        other["r101_bs8"], s_ = fwd_point(syn.R101_LEVELS, 8, 900, n_sets=2, steps=20)
        del s_
        torch.cuda.empty_cache()
        # headline shape: warm L2, backward, reference op
        dtype = torch.float32 if args.dtype == "f32" else torch.bfloat16
        sets = [to_device(d, dtype) for d in host]
        fwd = [(lambda g=g: cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"])) for g in sets]
        roof["fwd_warm_l2_us"] = time_graph(fwd[:1], 100, 10, True, nosync) / 100 * 1e3
        gf = torch.empty_like(sets[0]["feat"], dtype=torch.float32)
        bwd = [(lambda g=g: cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"], gf,
                                          torch.empty_like(g["loc"]), torch.empty_like(g["w"]),
                                          flags=cabi.BWD_OVERWRITE_SMALL | cabi.BWD_ZERO_GRAD_FEAT))
               for g in sets]
        roof["bwd_cold_us"] = time_graph(bwd, 40, 10, True, nosync) / 40 * 1e3
        bwd_nz = [(lambda g=g: cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"], gf,
                                             torch.empty_like(g["loc"]), torch.empty_like(g["w"]),
                                             flags=cabi.BWD_OVERWRITE_SMALL)) for g in sets]
        roof["bwd_cold_no_memset_us"] = time_graph(bwd_nz, 40, 10, True, nosync) / 40 * 1e3
        if ref is not None and args.dtype == "f32":
            rf = [(lambda g=g: ref.deformable_aggregation_forward(g["feat"], g["shape"], g["start"],
                                                                  g["loc"], g["w"])) for g in sets]
            g = sets[0]
            gfr, glr, gwr = torch.zeros_like(g["feat"]), torch.zeros_like(g["loc"]), torch.zeros_like(g["w"])
            rb = [(lambda g=g: (gfr.zero_(), glr.zero_(), gwr.zero_(),
                                ref.deformable_aggregation_backward(g["feat"], g["shape"], g["start"],
                                                                    g["loc"], g["w"], g["go"], gfr, glr, gwr)))
                  for g in sets]
            roof["reference_cuda_op"] = {
                "what": "the unmodified reference op (ops/src/deformable_aggregation_cuda.cu) rebuilt for sm_100a, "
                        "same GPU, same inputs, cold L2",
                "fwd_us": time_graph(rf, 40, 5, False, nosync) / 40 * 1e3,
                "bwd_us_incl_3_memsets": time_graph(rb, 20, 3, False, nosync) / 20 * 1e3}
    except Exception as e:  # pragma: no cover
        other["error"] = repr(e)
    roof["other_shapes"] = other
    # ---- BASELINE.json config #2: the SimPB+ R50 frame (frames/s), tools/frame_bench.py
    try:
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import frame_bench
        fr = frame_bench.run(frames=8, warmup=3)
        fr["whole_frame_cuda_graph"] = frame_bench.run_graph(frames=30, warmup=4)
        fr["whole_frame_cuda_graph_folded_bn"] = frame_bench.run_graph(frames=30, warmup=4, fold_bn=True)
        roof["frame"] = fr
    except Exception as e:  # pragma: no cover
        roof["frame"] = {"error": repr(e)}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
