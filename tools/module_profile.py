"""Kernel-level time table of one DeformableFeatureAggregation forward (released config, bs=1, eval)
with the torch profiler (CUDA activities).   python tools/module_profile.py"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpb_b200 import blocks, feature_maps_format, synthetic  # noqa: E402

torch.manual_seed(0)
m = blocks.DeformableFeatureAggregation(
    embed_dims=256, num_groups=8, num_levels=4, num_cams=6, attn_drop=0.15, use_deformable_func=True,
    use_camera_embed=True, residual_mode="cat",
    kps_generator=dict(type="SparseBox3DKeyPointsGenerator", num_learnable_pts=6,
                       fix_scale=synthetic.FIX_SCALE)).cuda().eval()
d = synthetic.module_inputs_rig(bs=1, seed=0)
g = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
fm = feature_maps_format([x.cuda() for x in d["feature_maps"]])
metas = dict(projection_mat=g["projection_mat"], image_wh=g["image_wh"])
with torch.no_grad():
    for _ in range(5):
        m(g["instance_feature"], g["anchor"], g["anchor_embed"], fm, metas)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(10):
            m(g["instance_feature"], g["anchor"], g["anchor_embed"], fm, metas)
        torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print("device time per forward: %.1f us" % (tot / 10))
for e in rows[:25]:
    print("%8.1f us  x%-3d %s" % (e.device_time_total / 10, e.count // 10, e.key[:110]))
