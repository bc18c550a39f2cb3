"""Tiny driver for ncu: a few cold-L2 launches of the forward (and optionally backward) kernel.
    python tools/profile_op.py [--batch B] [--inputs rig|uniform] [--dtype f32|bf16] [--bwd] [--iters N]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpb_b200 import cabi, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--anchors", type=int, default=900)
ap.add_argument("--inputs", default="rig")
ap.add_argument("--dtype", default="f32")
ap.add_argument("--bwd", action="store_true")
ap.add_argument("--iters", type=int, default=6)
a = ap.parse_args()
maker = synthetic.rig_op_inputs if a.inputs == "rig" else synthetic.op_inputs_uniform
dt = torch.float32 if a.dtype == "f32" else torch.bfloat16
n_sets = 4 if a.batch == 1 else 2
sets = []
for s in range(n_sets):
    d = maker(bs=a.batch, A=a.anchors, seed=s)
    sets.append(dict(feat=d["mc_ms_feat"].cuda().to(dt), shape=d["spatial_shape"].int().cuda(),
                     start=d["scale_start_index"].int().cuda(), loc=d["sampling_location"].cuda(),
                     w=d["weights"].cuda(), go=d["grad_output"].cuda()))
gf = torch.empty_like(sets[0]["feat"], dtype=torch.float32)
for i in range(a.iters):
    g = sets[i % n_sets]
    out = cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"])
    if a.bwd:
        cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"], gf,
                      torch.empty_like(g["loc"]), torch.empty_like(g["w"]),
                      flags=cabi.BWD_OVERWRITE_SMALL | cabi.BWD_ZERO_GRAD_FEAT)
torch.cuda.synchronize()
print("done", float(out.abs().sum()))
