"""Where does a CTA of the SM-pooled forward kernel spend its life (per batch iteration)?  Builds a tool-only copy of the
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
ap.add_argument("--variant", type=int, default=20)
ap.add_argument("--nw", type=int, default=4)
ap.add_argument("--inputs", default="rig")
a = ap.parse_args()
os.environ["DFA_FWD_VARIANT"] = str(a.variant)
lib_path = os.path.join(ROOT, "gpurun_out", "libdfa_b200_prof.so")
os.makedirs(os.path.dirname(lib_path), exist_ok=True)
objs, _ = build.compile_objects(extra_flags=["-DDFA_PHASE_TIMING"],
                                obj_dir=os.path.join(ROOT, "gpurun_out", "prof_objs"))
build.link_lib(objs, lib_path)
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
n_cta = 148 * 2
if a.variant == 21:
    pass
buf = torch.zeros(n_cta, 16, 16, dtype=torch.int64, device="cuda")
out = torch.empty(a.batch, 900, 256, device="cuda")
vp = ctypes.c_void_p
lib.dfa_debug_set_phase_buffer.argtypes = [vp]
lib.dfa_forward.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.POINTER(Dims), vp]
assert lib.dfa_debug_set_phase_buffer(buf.data_ptr()) == 0
st = torch.cuda.current_stream().cuda_stream
for i in range(3):   # the last launch (cold inputs: 3 sets rotate) is the one analysed
    g = sets[i]
    buf.zero_()
    dm = Dims(a.batch, 6, g["nf"], 256, 4, 900, 13, 8)
    rc = lib.dfa_forward(g["feat"].data_ptr(), 0, g["shape"].data_ptr(), g["start"].data_ptr(),
                         g["loc"].data_ptr(), g["w"].data_ptr(), out.data_ptr(), ctypes.byref(dm), st)
    assert rc == 0, rc
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
rc = lib.dfa_forward(g["feat"].data_ptr(), 0, g["shape"].data_ptr(), g["start"].data_ptr(),
                     g["loc"].data_ptr(), g["w"].data_ptr(), out.data_ptr(), ctypes.byref(dm), st)
e1.record()
torch.cuda.synchronize()
print("event time of one more launch (stamped build): %.1f us" % (e0.elapsed_time(e1) * 1e3))
t = buf.cpu().double()
gt = buf.cpu()[:, 15, :2]
gl = gt[:, 0] > 0
if gl.any():
    print("wall clock (globaltimer): CTA starts span %.2f us, first start to last end %.2f us, CTA life median %.2f / max %.2f us"
          % ((gt[gl, 0].max() - gt[gl, 0].min()) / 1e3, (gt[gl, 1].max() - gt[gl, 0].min()) / 1e3,
             float((gt[gl, 1] - gt[gl, 0]).double().median()) / 1e3, float((gt[gl, 1] - gt[gl, 0]).max()) / 1e3))
names = {0: "iteration top", 10: "sched: next copy issued", 11: "loc landed (warp 0)",
         1: "compaction barrier", 2: "records barrier", 4: "gather done (warp 0)",
         5: "gather barrier", 6: "reduce done (thread 0)"}
start = t[:, 0, 0]
live = start > 0
print("variant", a.variant, "batch", a.batch, "CTAs", int(live.sum()))
for it in range(8):
    ok = live & (t[:, it, 0] > 0)
    if not ok.any():
        break
    print("iteration %d (%d CTAs): cycles since CTA start, median / p90 / max" % (it, int(ok.sum())))
    for i in (0, 11, 1, 2, 10, 4, 5, 6):
        x = (t[:, it, i] - start)[ok & (t[:, it, i] > 0)]
        if x.numel():
            print("  %-28s %8.0f %8.0f %8.0f" % (names[i], x.median(), x.quantile(0.9), x.max()))
