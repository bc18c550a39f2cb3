"""Back-of-the-envelope model of the forward's tail on a one-wave grid: a CTA lives
prologue + rounds x round time + epilogue, with rounds = valid samples of its anchor (four slices, four
levels: one round per valid sample).  Constants are the medians of profiles/r1_fwd_rows_timeline_bs1.txt.
Prints the time of the last CTA for the measured sample-count distribution when groups of m anchors
pool their taps over m CTAs (m = 1 is the shipped kernel) — the most any work-sharing scheme inside a
group of that size can gain.   python tools/tail_model.py        (CPU only)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpb_b200 import synthetic  # noqa: E402

PROLOGUE, ROUND, EPILOGUE, GHZ = 4938.0, 1100.0, 1118.0, 1.965      # cycles, cycles per round, cycles, GHz
rows = []
for seed in range(5):
    d = synthetic.rig_op_inputs(bs=1, seed=seed, feat=False)
    loc = d["sampling_location"]
    nv = ((loc > 0) & (loc < 1)).all(-1).flatten(2).sum(-1).flatten().double()[:888]      # first wave
    res = []
    for m in (1, 2, 4, 8, 888):
        grp = nv[: (888 // m) * m].view(-1, m).sum(1) / m                                   # rounds per CTA of a group
        res.append(float((PROLOGUE + torch.ceil(grp).max() * ROUND + EPILOGUE) / GHZ / 1e3))
    rows.append(res)
    print("seed %d: valid samples median %d p90 %d max %d | last CTA ends after (us): m=1 %.1f  m=2 %.1f  m=4 %.1f  "
          "m=8 %.1f  perfectly balanced %.1f" % (seed, nv.median(), nv.quantile(0.9), nv.max(), *res))
t = torch.tensor(rows).mean(0)
print("mean: m=1 %.1f  m=2 %.1f  m=4 %.1f  m=8 %.1f  balanced %.1f us  (measured, m=1: heavy anchors end at 16.6-16.9 us)"
      % tuple(t.tolist()))
