"""How much would spatially sorted anchors help?  Probe, not product: the anchors of each sample are
sorted by (camera of their centre point, coarse image cell) on the host and dealt to CTAs so that an
SM (round-robin CTA dispatch assumed) sees neighbours back to back; the unchanged forward kernel is
timed on the original and on the reordered inputs (cold L2).
    python tools/locality_probe.py [--batch B]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from simpb_b200 import cabi, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
a = ap.parse_args()
n_sets = 2 if a.batch > 1 else 4
host = [synthetic.rig_op_inputs(bs=a.batch, seed=s) for s in range(n_sets)]


def reorder(d, interleave):
    loc, w = d["sampling_location"], d["weights"]
    bs, A = loc.shape[:2]
    c = loc[:, :, 0]                                   # centre key point: [bs, A, K, 2]
    valid = ((c > 0) & (c < 1)).all(-1)                # [bs, A, K]
    cam = torch.where(valid.any(-1), valid.float().argmax(-1), torch.full((bs, A), 6))
    xy = torch.gather(c, 2, cam.clamp(max=5)[..., None, None].expand(bs, A, 1, 2))[:, :, 0]
    cell = (xy[..., 1].clamp(0, 0.999) * 8).long() * 22 + (xy[..., 0].clamp(0, 0.999) * 22).long()
    key = cam * 1000 + cell
    order = key.argsort(dim=1)
    if interleave:                                     # CTA c -> anchor (c % 148) * chunk + c // 148
        chunk = (A + 147) // 148
        cta = torch.arange(148 * chunk)
        src = (cta % 148) * chunk + cta // 148
        src = src[src < A]
        order = order[:, src]
    idx = order[..., None, None, None].expand_as(loc)
    out = dict(d)
    out["sampling_location"] = torch.gather(loc, 1, idx).contiguous()
    out["weights"] = torch.gather(w, 1, order[..., None, None, None, None].expand_as(w)).contiguous()
    return out


def timeit(sets):
    dev = [bench.to_device(d, torch.float32) for d in sets]
    outs = [torch.empty(a.batch, 900, 256, device="cuda") for _ in dev]
    fns = [(lambda g=g, o=o: cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], out=o))
           for g, o in zip(dev, outs)]
    return bench.time_graph(fns, 100, 10, True, torch.cuda.synchronize) / 100 * 1e3


print("batch %d" % a.batch)
print("  original order          %.2f us" % timeit(host))
print("  sorted (blockIdx order)  %.2f us" % timeit([reorder(d, False) for d in host]))
print("  sorted + per-SM chunks   %.2f us" % timeit([reorder(d, True) for d in host]))
