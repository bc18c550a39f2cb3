"""Small end-to-end pass over every kernel family for compute-sanitizer (one tool per run):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpb_b200 import blocks, cabi, feature_maps_format, synthetic  # noqa: E402

LV = ((16, 44), (8, 22), (4, 11), (2, 6))


def dev(d, dt=torch.float32):
    return (d["mc_ms_feat"].cuda().to(dt), d["spatial_shape"].int().cuda(), d["scale_start_index"].int().cuda(),
            d["sampling_location"].cuda(), d["weights"].cuda(), d["grad_output"].cuda())


cases = [synthetic.rig_op_inputs(bs=2, A=40, levels=LV, seed=0),
         synthetic.op_inputs_uniform(bs=1, A=12, levels=LV, seed=1),
         synthetic.op_inputs_uniform(bs=1, A=5, P=3, K=3, levels=LV[:3], C=256, G=8, seed=2),   # no TMA
         synthetic.op_inputs_uniform(bs=1, A=4, P=40, K=6, levels=LV, seed=3),                  # many samples
         synthetic.op_inputs_uniform(bs=1, A=6, P=5, K=2, levels=LV[:2], C=32, G=4, seed=4),    # per-group kernels
         synthetic.op_inputs_uniform(bs=1, A=6, P=3, K=1, levels=LV[:1], C=6, G=3, seed=5)]     # generic kernels
for fv in ("1", "2", "3", "0", "30", "33"):
    os.environ["DFA_FWD_VARIANT"] = fv
    cabi.reload_knobs()
    for d in cases:
        for dt in (torch.float32, torch.bfloat16):
            f, sh, st, loc, w, go = dev(d, dt)
            cabi.forward(f, sh, st, loc, w)
for bv in ("10", "11", "12", "0"):
    os.environ["DFA_BWD_VARIANT"] = bv
    cabi.reload_knobs()
    for d in cases:
        for dt in (torch.float32, torch.bfloat16):
            f, sh, st, loc, w, go = dev(d, dt)
            cabi.backward(f, sh, st, loc, w, go)
            cabi.backward(f, sh, st, loc, w, go, need_feat=False)
# host-buffer entry point: pull mode (pinned, mapped buffers) and whole copies (pageable)
for d in cases[:3]:
    for dt in (torch.float32, torch.bfloat16):
        bs, A, P, K = d["sampling_location"].shape[:4]
        L, G = d["weights"].shape[4:6]
        C = d["mc_ms_feat"].shape[2]
        hf = cabi.HostForward(cabi.Dims(bs, K, d["num_feat"], C, L, A, P, G), dt)
        host = [d["mc_ms_feat"].to(dt).contiguous(), d["spatial_shape"].int().contiguous(),
                d["scale_start_index"].int().contiguous(), d["sampling_location"].contiguous(), d["weights"].contiguous()]
        a = hf(*[t.pin_memory() for t in host], torch.empty(bs, A, C).pin_memory()).clone()
        b = hf(*host, torch.empty(bs, A, C))
        assert torch.equal(a, b), "pull mode differs from whole copies"
        hf.stats()
torch.cuda.synchronize()
print("op kernels done")

# front end + module
m = blocks.DeformableFeatureAggregation(embed_dims=256, num_groups=8, num_levels=4, num_cams=6, attn_drop=0.15,
                                        use_camera_embed=True, residual_mode="cat",
                                        kps_generator=dict(type="SparseBox3DKeyPointsGenerator", num_learnable_pts=6,
                                                           fix_scale=synthetic.FIX_SCALE)).cuda().train()
d = synthetic.module_inputs_rig(bs=2, A=30, levels=LV, seed=6)
g = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
maps = [x.cuda().requires_grad_() for x in d["feature_maps"]]
inst = g["instance_feature"].requires_grad_()
out = m(inst, g["anchor"].requires_grad_(), g["anchor_embed"], feature_maps_format(maps),
        dict(projection_mat=g["projection_mat"], image_wh=g["image_wh"]))
out.sum().backward()
m2 = blocks.DeformableFeatureAggregation(embed_dims=32, num_groups=2, num_levels=4, num_cams=6,
                                         kps_generator=dict(type="SparseBox3DKeyPointsGenerator")).cuda()
maps2 = [torch.randn(1, 6, 32, h, w).cuda() for h, w in LV]
m2(torch.randn(1, 7, 32).cuda(), g["anchor"][:1, :7].detach(), torch.randn(1, 7, 32).cuda(),
   feature_maps_format(maps2, dtype=torch.bfloat16), dict(projection_mat=g["projection_mat"][:1])).sum().backward()
torch.cuda.synchronize()
print("module done")

# MSDA
shapes = torch.tensor(LV, dtype=torch.int32).cuda()
cnt = (shapes[:, 0] * shapes[:, 1]).long()
start = torch.cat([cnt.new_zeros(1), cnt.cumsum(0)[:-1]]).int()
S = int(cnt.sum())
for (M, D, P, K) in ((8, 32, 4, 3), (4, 8, 3, 1), (3, 4, 2, 1), (2, 5, 3, 1)):
    for dt in (torch.float32, torch.bfloat16):
        val = torch.randn(2, K, S, M, D).cuda().to(dt) if K > 1 else torch.randn(2, S, M, D).cuda().to(dt)
        loc = (torch.rand(2, 9, M, 4, P, 2) * 1.3 - 0.15).cuda()
        w = torch.rand(2, 9, M, 4, P).cuda()
        table = torch.randint(0, K, (9,), dtype=torch.int32).cuda() if K > 1 else None
        cabi.msda_forward(val, shapes, start, loc, w, table)
        cabi.msda_backward(val, shapes, start, loc, w, torch.randn(2, 9, M * D).cuda(), query_table=table)
torch.cuda.synchronize()
print("msda done")
