"""Kernel-time table of one SimPB+ R50 frame (torch.profiler, CUDA activities): which kernels the
13 ms of a graph-replayed frame consist of.   python tools/frame_profile.py [--static 320]"""
import argparse
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpb_b200 import decoder, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--static", type=int, default=320)
ap.add_argument("--top", type=int, default=40)
a = ap.parse_args()
dev = "cuda"
model = decoder.SimPBFrame(seed=0, static_queries=a.static or None).to(dev).eval()
proj, wh = synthetic.camera_rig(1)
T = torch.eye(4)[None].clone()
T[0, 1, 3] = -2.5
metas = dict(projection_mat=proj.to(dev), image_wh=wh.to(dev), img_wh=(704.0, 256.0), T_temp2cur=T.to(dev),
             dt=torch.full((1,), 0.5, device=dev))
img = torch.randn(1, 6, 3, 256, 704, device=dev)
with torch.no_grad():
    for _ in range(3):
        model(img, metas)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        model(img, metas)
        torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        tot[e.name][0] += e.device_time if hasattr(e, "device_time") else e.cuda_time
        tot[e.name][1] += 1
rows = sorted(tot.items(), key=lambda kv: -kv[1][0])
total = sum(v[0] for _, v in rows)
print("device time of one frame: %.3f ms in %d kernels" % (total / 1e3, sum(v[1] for _, v in rows)))
for name, (us, n) in rows[:a.top]:
    print("%9.1f us %5.1f%% %5d  %s" % (us, 100 * us / total, n, name[:150]))
