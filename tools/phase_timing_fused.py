"""clock64 phase stamps of the fused module forward kernel (tool-only build, -DDFA_PHASE_TIMING).
    python tools/phase_timing_fused.py [--batch B]"""
import argparse
import ctypes
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("simpb_b200_build", os.path.join(ROOT, "simpb_b200", "build.py"))
build = importlib.util.module_from_spec(spec)
spec.loader.exec_module(build)
from simpb_b200 import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
a = ap.parse_args()
lib_path = os.path.join(ROOT, "gpurun_out", "libdfa_b200_prof.so")
objs, _ = build.compile_objects(extra_flags=["-DDFA_PHASE_TIMING"], obj_dir=os.path.join(ROOT, "gpurun_out", "prof_objs"))
build.link_lib(objs, lib_path)
lib = ctypes.CDLL(lib_path)


class Dims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("bs", "K", "nf", "C", "L", "A", "P", "G")]


bs, A, K, L, P, G = a.batch, 900, 6, 4, 13, 8
vp = ctypes.c_void_p
lib.dfa_debug_set_phase_buffer.argtypes = [vp]
lib.dfa_forward_fused.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp,
                                  ctypes.POINTER(Dims), vp]
buf = torch.zeros(bs * A, 8, 8, dtype=torch.int64, device="cuda")
assert lib.dfa_debug_set_phase_buffer(buf.data_ptr()) == 0
out = torch.empty(bs, A, 256, device="cuda")
fix = torch.tensor(synthetic.FIX_SCALE).cuda()
for s in range(3):
    d = synthetic.module_inputs_rig(bs=bs, seed=s, feat=False)
    gen = torch.Generator().manual_seed(s)
    shape, start, nf = synthetic.level_tables()
    feat = torch.randn(bs, nf, 256, generator=gen).cuda()
    la = torch.randn(bs, A, L * P * G, generator=gen).cuda()
    lk = torch.randn(bs, K, L * P * G, generator=gen).cuda()
    off = torch.randn(bs, A, 18, generator=gen).cuda()
    anchor, proj, wh = d["anchor"].cuda(), d["projection_mat"].cuda(), d["image_wh"].cuda()
    sh, st = shape.int().cuda(), start.int().cuda()
    dm = Dims(bs, K, nf, 256, L, A, P, G)
    torch.cuda.synchronize()
    rc = lib.dfa_forward_fused(feat.data_ptr(), 0, sh.data_ptr(), st.data_ptr(), anchor.data_ptr(), fix.data_ptr(), 7,
                               off.data_ptr(), proj.data_ptr(), wh.data_ptr(), la.data_ptr(), lk.data_ptr(),
                               out.data_ptr(), None, ctypes.byref(dm), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
torch.cuda.synchronize()
t = buf.cpu().double()
names = ["0 start", "1 key points", "2 projections", "3 mask + logits landed", "4 softmax numerators", "5 tap records",
         "6 gather", "7 end"]
t0 = t[:, :, 0].min(dim=1, keepdim=True).values
print("batch", bs, "- stamps relative to CTA start, cycles: median / p90 over (anchor, warp)")
for i in range(1, 8):
    x = (t[:, :, i] - t0).flatten()
    print("  %-26s %8.0f %8.0f" % (names[i], x.median(), x.quantile(0.9)))
