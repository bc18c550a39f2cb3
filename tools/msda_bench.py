"""Multi-scale deformable attention at SimPB's 2-D branch shape: 6 camera groups over the R50
704x256 pyramid (14,960 rows x 8 heads x 32 channels per camera), ~1000 2-D queries per sample split
over the cameras, 4 levels x 4 points.  One grouped launch vs the reference's loop of per-group calls
(same kernels), forward and backward, cold L2 (rotating value tables).
    python tools/msda_bench.py [--batch B] [--queries Q]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from simpb_b200 import cabi, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--queries", type=int, default=1008)
ap.add_argument("--dtype", default="f32")
a = ap.parse_args()
K, M, D, L, P = 6, 8, 32, 4, 4
shapes = torch.tensor(synthetic.R50_LEVELS, dtype=torch.int32)
counts = (shapes[:, 0] * shapes[:, 1]).long()
start = torch.cat([counts.new_zeros(1), counts.cumsum(0)[:-1]]).int()
S = int(counts.sum())
per = a.queries // K
groups = [(i * per, (i + 1) * per) for i in range(K)]
table = torch.repeat_interleave(torch.arange(K, dtype=torch.int32), per).cuda()
Q = per * K
dt = torch.float32 if a.dtype == "f32" else torch.bfloat16
gen = torch.Generator().manual_seed(0)
sets = []
for s in range(4 if a.batch == 1 else 2):
    loc = torch.rand(a.batch, Q, 1, 1, 1, 2, generator=gen) + 0.05 * torch.randn(a.batch, Q, M, L, P, 2, generator=gen)
    sets.append(dict(value=torch.randn(a.batch, K, S, M, D, generator=gen).cuda().to(dt),
                     loc=loc.cuda().contiguous(),
                     w=torch.rand(a.batch, Q, M, L * P, generator=gen).softmax(-1).view(a.batch, Q, M, L, P).cuda(),
                     go=torch.randn(a.batch, Q, M * D, generator=gen).cuda()))
sh, st = shapes.cuda(), start.cuda()
sync = torch.cuda.synchronize
fwd = [(lambda g=g: cabi.msda_forward(g["value"], sh, st, g["loc"], g["w"], table)) for g in sets]
bwd = [(lambda g=g: cabi.msda_backward(g["value"], sh, st, g["loc"], g["w"], g["go"], query_table=table)) for g in sets]


def looped(g):
    for i, (q0, q1) in enumerate(groups):
        cabi.msda_forward(g["value"][:, i], sh, st, g["loc"][:, q0:q1], g["w"][:, q0:q1])


for g in sets:      # per-group views must be contiguous for the loop arm
    g["value_i"] = [g["value"][:, i].contiguous() for i in range(K)]
    g["loc_i"] = [g["loc"][:, q0:q1].contiguous() for q0, q1 in groups]
    g["w_i"] = [g["w"][:, q0:q1].contiguous() for q0, q1 in groups]
loop = [(lambda g=g: [cabi.msda_forward(g["value_i"][i], sh, st, g["loc_i"][i], g["w_i"][i]) for i in range(K)])
        for g in sets]
t_f = bench.time_graph(fwd, 80, 8, True, sync) / 80
t_l = bench.time_graph(loop, 80, 8, True, sync) / 80
t_b = bench.time_graph(bwd, 40, 4, True, sync) / 40
esz = 4 if a.dtype == "f32" else 2
gathered = a.batch * Q * M * L * P * 4 * D * esz           # bytes requested from the value tables
small = a.batch * Q * M * L * P * 12 + a.batch * Q * M * D * 4
print(json.dumps({"config": "MSDA, %d queries in %d camera groups, bs=%d, R50 pyramid per camera, %s" % (Q, K, a.batch, a.dtype),
                  "fwd_grouped_us": round(t_f * 1e3, 2), "fwd_per_group_loop_us": round(t_l * 1e3, 2),
                  "bwd_grouped_us_incl_fill": round(t_b * 1e3, 2),
                  "queries_per_s": round(a.batch * Q / (t_f * 1e-3)),
                  "gathered_MB": round(gathered / 1e6, 1), "loc_w_out_MB": round(small / 1e6, 2),
                  "gathered_GBps": round((gathered + small) / (t_f * 1e-3) / 1e9)}))
