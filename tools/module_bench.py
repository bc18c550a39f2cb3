"""DeformableFeatureAggregation module forward at the released config (SimPB+ R50 704x256, bs=1,
eval): the fused front end of simpb_b200/blocks.py against the reference's PyTorch front end
(restated in oracle/module_ref.py, run on the GPU) feeding the same CUDA op.  Three DFA layers per
frame, as in the released decoder.  Eager and CUDA-graph timings, cold feature maps.
    python tools/module_bench.py [--batch B]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import module_ref  # noqa: E402
from simpb_b200 import blocks, deformable_aggregation_function, feature_maps_format, msda, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--train", action="store_true", help="time forward + backward of ONE layer in training mode")
a = ap.parse_args()
torch.manual_seed(0)
cfg = dict(embed_dims=256, num_groups=8, num_levels=4, num_cams=6, attn_drop=0.15,
           use_deformable_func=True, use_camera_embed=True, residual_mode="cat",
           kps_generator=dict(type="SparseBox3DKeyPointsGenerator", num_learnable_pts=6,
                              fix_scale=synthetic.FIX_SCALE))
layers = [blocks.DeformableFeatureAggregation(**cfg).cuda().eval() for _ in range(3)]
refs = []
for m in layers:
    r = module_ref.DFAModuleRef(256, 8, 4, 6, attn_drop=0.15, fix_scale=synthetic.FIX_SCALE,
                                num_learnable_pts=6, use_camera_embed=True, residual_mode="cat",
                                op=lambda col, sh, st, loc, w: deformable_aggregation_function(col, sh, st, loc, w))
    r.load_state_dict(m.state_dict())
    refs.append(r.cuda().eval())

frames = []
for s in range(3):           # rotating frames: > L2 of feature maps in total
    d = synthetic.module_inputs_rig(bs=a.batch, seed=s)
    g = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    g["fm"] = feature_maps_format([m.cuda() for m in d["feature_maps"]])
    g["metas"] = dict(projection_mat=g["projection_mat"], image_wh=g["image_wh"])
    frames.append(g)


def run_fused(g):
    x = g["instance_feature"]
    for m in layers:
        x = m(x, g["anchor"], g["anchor_embed"], g["fm"], g["metas"])[..., :256]
    return x


def run_ref(g):
    x = g["instance_feature"]
    col, sh, st = g["fm"]
    for r in refs:
        logits = r.kps_generator.learnable_fc(x)
        pts = module_ref.key_points(g["anchor"], r.kps_generator.fix_scale, logits)
        w = r.attention_weights(x, g["anchor_embed"], g["projection_mat"])
        uv = module_ref.project_points(pts, g["projection_mat"], g["image_wh"])
        feats = deformable_aggregation_function(col, sh, st, uv.permute(0, 2, 3, 1, 4).contiguous(),
                                                w.permute(0, 1, 4, 2, 3, 5).contiguous())
        x = r.output_proj(feats)
    return x


def train_step(front, g):
    """forward + backward of one DFA layer with gradients for inputs, feature table and parameters."""
    inst = g["instance_feature"].detach().requires_grad_()
    col = g["fm"][0].detach().requires_grad_()
    if front == "fused":
        m = layers[0]
        out = m(inst, g["anchor"], g["anchor_embed"], [col, g["fm"][1], g["fm"][2]], g["metas"])
    else:
        r = refs[0]
        logits = r.kps_generator.learnable_fc(inst)
        pts = module_ref.key_points(g["anchor"], r.kps_generator.fix_scale, logits)
        keep = (torch.rand(a.batch, 900, 6, 1, 13, 1, device="cuda") > 0.15)
        w = r.attention_weights(inst, g["anchor_embed"], g["projection_mat"], drop_mask=keep)
        uv = module_ref.project_points(pts, g["projection_mat"], g["image_wh"])
        feats = deformable_aggregation_function(col, g["fm"][1], g["fm"][2], uv.permute(0, 2, 3, 1, 4).contiguous(),
                                                w.permute(0, 1, 4, 2, 3, 5).contiguous())
        out = torch.cat([r.output_proj(feats), inst], -1)
    out.sum().backward()


if a.train:
    for m in layers + refs:
        m.train()
    res = {"config": "1 x DFA layer forward + backward, training mode (attn-drop 0.15), bs=%d, 900 anchors" % a.batch}
    for front in ("fused", "torch_front_end"):
        for _ in range(3):
            for g in frames:
                train_step(front, g)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            for g in frames:
                train_step(front, g)
        e1.record()
        torch.cuda.synchronize()
        res[front + "_eager_ms_per_step"] = round(e0.elapsed_time(e1) / (10 * len(frames)), 4)
    print(json.dumps(res))
    sys.exit(0)


def timeit(fn, graph):
    with torch.no_grad():
        for g in frames:
            fn(g)
        torch.cuda.synchronize()
        gr = None
        if graph:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for g in frames:
                    fn(g)
            torch.cuda.current_stream().wait_stream(s)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for g in frames:
                    fn(g)
            gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            if gr is not None:
                gr.replay()
            else:
                for g in frames:
                    fn(g)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(frames))


# the 2-D branch of the released decoder: 3 x QueryGroupMultiScaleDeformableAttention on ~1000 2-D
# queries split over the 6 cameras, reading the same channel-last table
msda_layers = [msda.QueryGroupMultiScaleDeformableAttention(256, 8, 4, 4, 6, dropout=0.0, batch_first=True,
                                                            residual_mode="cat").cuda().eval() for _ in range(3)]
Q2 = 1008
groups = [(i * Q2 // 6, (i + 1) * Q2 // 6) for i in range(6)]
gen = torch.Generator().manual_seed(5)
q2d = torch.randn(a.batch, Q2, 256, generator=gen).cuda()
ref2d = torch.rand(a.batch, Q2, 4, 2, generator=gen).cuda()


def run_both(g):
    x = run_fused(g)
    col, sh, st = g["fm"]
    value = col.reshape(a.batch, 6, -1, 256).flatten(0, 1)            # simpb_head.py:282-292
    y = q2d
    for m in msda_layers:
        y = m(y, value=value, reference_points=ref2d, spatial_shapes=sh[0], level_start_index=st[0],
              query_groups=groups)[..., :256]
    return x, y


with torch.no_grad():
    err = float((run_fused(frames[0]) - run_ref(frames[0])).abs().max() / run_ref(frames[0]).abs().max())
out = {"config": "3 x DFA (released SimPB+ R50 config), bs=%d, 900 anchors, eval" % a.batch,
       "fused_vs_torch_front_end_max_rel_diff": err}
for name, fn in (("fused", run_fused), ("torch_front_end", run_ref)):
    for graph in (False, True):
        key = "%s_%s_ms_per_frame" % (name, "graph" if graph else "eager")
        try:
            out[key] = round(timeit(fn, graph), 4)
        except RuntimeError as e:     # the reference front end indexes with Python lists (a CPU index
            out[key] = None           # tensor per call): it cannot be captured into a CUDA graph
            out[key + "_error"] = str(e).split(".")[0][:120]
            torch.cuda.synchronize()
out["dfa_plus_msda_graph_ms_per_frame"] = round(timeit(run_both, True), 4)
out["dfa_plus_msda_frames_per_s"] = round(1e3 / out["dfa_plus_msda_graph_ms_per_frame"], 1)
print(json.dumps(out))
