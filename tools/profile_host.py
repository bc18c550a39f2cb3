"""Tiny driver for ncu: a few calls of dfa_forward_host in pull mode (pinned host buffers, R50 rig inputs)
and of the MSDA module's gather-then-project kernel.   python tools/profile_host.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpb_b200 import cabi, synthetic  # noqa: E402

d = synthetic.rig_op_inputs(bs=1, seed=0)
dims = cabi.Dims(1, 6, d["num_feat"], 256, 4, 900, 13, 8)
hf = cabi.HostForward(dims)
pin = lambda t: t.contiguous().pin_memory()  # noqa: E731
h = [pin(d["mc_ms_feat"]), pin(d["spatial_shape"].int()), pin(d["scale_start_index"].int()),
     pin(d["sampling_location"]), pin(d["weights"])]
out = torch.empty(1, 900, 256).pin_memory()
for _ in range(4):
    hf(*h, out)
print("pull mode moved", hf.stats())

# MSDA on the unprojected table: 1920 queries in 6 camera groups, R50 pyramid per camera
levels = synthetic.R50_LEVELS
shapes = torch.tensor(levels, dtype=torch.int32).cuda()
cnt = (shapes[:, 0] * shapes[:, 1]).long()
start = torch.cat([cnt.new_zeros(1), cnt.cumsum(0)[:-1]]).int()
S = int(cnt.sum())
table = torch.randn(1, 6, S, 256, device="cuda")
loc = torch.rand(1, 1920, 8, 4, 4, 2, device="cuda")
w = torch.rand(1, 1920, 8, 16, device="cuda").softmax(-1).view(1, 1920, 8, 4, 4)
qt = torch.arange(6, dtype=torch.int32, device="cuda").repeat_interleave(320)
for _ in range(4):
    g, s = cabi.msda_forward_raw(table, shapes, start, loc, w, qt)
torch.cuda.synchronize()
print("done", float(g.abs().sum()))
