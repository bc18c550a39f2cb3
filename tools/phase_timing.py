"""Where does a CTA of the merging forward kernel spend its life?  Builds a tool-only copy of the
library with -DDFA_PHASE_TIMING (clock64 stamps at phase boundaries) and prints per-phase cycle
statistics.   python tools/phase_timing.py [--batch B] [--variant V]
"""
import argparse
import ctypes
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simpb_b200 import build, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--variant", type=int, default=10)
ap.add_argument("--nw", type=int, default=4)
ap.add_argument("--inputs", default="rig")
ap.add_argument("--build-only", action="store_true")
a = ap.parse_args()
os.environ["DFA_FWD_VARIANT"] = str(a.variant)
prof_dir = os.path.join(ROOT, "tools", "_prof")          # git-ignored; travels with the snapshot, so build it
lib_path = os.path.join(prof_dir, "libdfa_b200_prof.so")   # in the container (python tools/phase_timing.py --build-only)
os.makedirs(prof_dir, exist_ok=True)
if build._stale(lib_path, build.KERNEL_SRCS + build.HEADERS + [build.HEADER]):
    objs, _ = build.compile_objects(extra_flags=["-DDFA_PHASE_TIMING"], obj_dir=os.path.join(prof_dir, "objs"))
    build.link_lib(objs, lib_path)
if a.build_only:
    sys.exit(0)
lib = ctypes.CDLL(lib_path)


class Dims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("bs", "K", "nf", "C", "L", "A", "P", "G")]


maker = synthetic.rig_op_inputs if a.inputs == "rig" else synthetic.op_inputs_uniform
sets = []
for s in range(3):
    d = maker(bs=a.batch, seed=s)
    sets.append(dict(feat=d["mc_ms_feat"].cuda(), shape=d["spatial_shape"].int().cuda(),
                     start=d["scale_start_index"].int().cuda(), loc=d["sampling_location"].cuda(),
                     w=d["weights"].cuda(), nf=d["num_feat"]))
n_anchor = a.batch * 900
buf = torch.zeros(n_anchor * 4, a.nw, 8, dtype=torch.int64, device="cuda")   # room for split CTAs
out = torch.empty(a.batch, 900, 256, device="cuda")
vp = ctypes.c_void_p
lib.dfa_debug_set_phase_buffer.argtypes = [vp]
lib.dfa_forward.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.POINTER(Dims), vp]
assert lib.dfa_debug_set_phase_buffer(buf.data_ptr()) == 0
st = torch.cuda.current_stream().cuda_stream
for i in range(3):   # the last launch (cold inputs: 3 sets rotate) is the one analysed
    g = sets[i]
    dm = Dims(a.batch, 6, g["nf"], 256, 4, 900, 13, 8)
    rc = lib.dfa_forward(g["feat"].data_ptr(), 0, g["shape"].data_ptr(), g["start"].data_ptr(),
                         g["loc"].data_ptr(), g["w"].data_ptr(), out.data_ptr(), ctypes.byref(dm), st)
    assert rc == 0, rc
torch.cuda.synchronize()
t = buf.cpu().double()
t = t[t[:, 0, 0] > 0]       # CTAs that ran
rows_like = a.variant < 10 or a.variant >= 30
if a.variant >= 40:    # channel-sliced kernel
    names = ["0 start", "1 locations landed, camera masks", "2 barrier + plan (+ weights landed)", "3 fine rows built, barrier",
             "4 gather done", "5 end", "-", "-"]
elif a.variant >= 30:    # window-merging kernel
    names = ["0 start", "1 locations landed, camera masks", "2 barrier (+ weights landed)", "3 fine-level taps done",
             "4 coarse-level units done", "5 end", "-", "-"]
elif a.variant < 10:     # row-sliced kernel: its own stamp meanings
    names = ["0 start", "1 operands staged + compaction", "2 tap records", "3 barrier + weights landed",
             "4 gather done", "5 end", "-", "-"]
else:
    names = ["0 start", "1 loc landed / tables", "2 compaction+weights issued", "3 barrier A", "4 merge done",
             "5 barrier B", "6 gather done", "7 end"]
if os.environ.get("PHASE_DEBUG"):
    torch.set_printoptions(precision=0, linewidth=200, sci_mode=False)
    print(t[0].long()); print(t[5].long()); print(t.shape)
t0 = t[:, :, 0].min(dim=1, keepdim=True).values          # CTA start
print("variant", a.variant, "batch", a.batch, "anchors", n_anchor)
print("phase stamps relative to CTA start, cycles: median / p90 / max over (anchor, warp)")
for i in range(1, 6 if rows_like else 8):
    x = (t[:, :, i] - t0).flatten()
    x = x[t[:, :, i].flatten() > 0]
    if x.numel():
        print("  %-30s %8.0f %8.0f %8.0f" % (names[i], x.median(), x.quantile(0.9), x.max()))
last = 5 if rows_like else 7
life = (t[:, :, last].max(dim=1).values - t0[:, 0])
print("CTA lifetime: median %.0f p90 %.0f max %.0f cycles" % (life.median(), life.quantile(0.9), life.max()))
if rows_like:     # globaltimer (ns) at CTA start / end: launch ramp and tail in wall time
    gs, ge = t[:, 0, 6], t[:, :, 7].max(dim=1).values
    first = gs.min()
    print("wall clock (globaltimer, us): CTA starts  median %.2f  p90 %.2f  max %.2f   |   CTA ends  median %.2f  p90 %.2f  "
          "p99 %.2f  max %.2f" % tuple(float(x) / 1e3 for x in (
              (gs - first).median(), (gs - first).quantile(0.9), (gs - first).max(), (ge - first).median(),
              (ge - first).quantile(0.9), (ge - first).quantile(0.99), (ge - first).max())))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = sets[0]
    lib.dfa_debug_set_phase_buffer(None)
    torch.cuda.synchronize()
    e0.record()
    lib.dfa_forward(g["feat"].data_ptr(), 0, g["shape"].data_ptr(), g["start"].data_ptr(), g["loc"].data_ptr(),
                    g["w"].data_ptr(), out.data_ptr(), ctypes.byref(dm), st)
    e1.record()
    torch.cuda.synchronize()
    print("event time of one isolated launch (no stamps written): %.2f us" % (e0.elapsed_time(e1) * 1e3))
    nvalid = ((sets[2]["loc"] > 0) & (sets[2]["loc"] < 1)).all(-1).flatten(2).sum(-1).flatten().cpu().double()
    gs, ge = gs[:nvalid.numel()], ge[:nvalid.numel()]
    nvalid = nvalid[:gs.numel()]
    lf = (ge - gs) / 1e3
    order = torch.argsort(lf, descending=True)[:8]
    print("longest-lived anchors: " + ", ".join("%d valid samples %.1f us (start +%.1f)" % (
        int(nvalid[i]), float(lf[i]), float(gs[i] - first) / 1e3) for i in order))
    print("valid samples per anchor: median %d  p90 %d  max %d;  corr(life, valid) = %.2f"
          % (nvalid.median(), nvalid.quantile(0.9), nvalid.max(), float(torch.corrcoef(torch.stack([lf, nvalid]))[0, 1])))
span = t[:, :, last].max() - t[:, :, 0][t[:, :, 0] > 0].min()
print("(clock64 is per SM; first start to last end across SMs is only indicative: %.0f cycles)" % span)
