"""Times the forward kernel variants (DFA_FWD_VARIANT) on cold-L2 rotating inputs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpb_b200 import cabi, synthetic  # noqa: E402
import bench  # noqa: E402

variants = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1,2,10,11,12").split(",")]
cases = [("rig", 1, "f32"), ("uniform", 1, "f32"), ("rig", 8, "f32"), ("rig", 1, "bf16"), ("rig", 8, "bf16")]
if len(sys.argv) > 2:      # e.g. "rig:1:f32,rig:8:f32"
    cases = [(c.split(":")[0], int(c.split(":")[1]), c.split(":")[2]) for c in sys.argv[2].split(",")]
for inputs, batch, dt in cases:
    maker = synthetic.rig_op_inputs if inputs == "rig" else synthetic.op_inputs_uniform
    dtype = torch.float32 if dt == "f32" else torch.bfloat16
    n_sets = 5 if batch == 1 else 2
    sets = [bench.to_device(maker(bs=batch, seed=s), dtype) for s in range(n_sets)]
    outs = [torch.empty(batch, 900, 256, device="cuda") for _ in sets]
    ref = None
    for v, pf in [(v, 0) for v in variants]:
        os.environ["DFA_FWD_VARIANT"] = str(v)
        cabi.reload_knobs()
        fns = [(lambda g=g, o=o: cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], out=o))
               for g, o in zip(sets, outs)]
        ms = bench.time_graph(fns, 200, 20, True, torch.cuda.synchronize) / 200
        cur = outs[0].clone()
        if ref is None:
            ref = cur
        err = float((cur - ref).abs().max() / ref.abs().max())
        print("%-8s bs=%d %-4s variant=%d pf=%d  %8.2f us   (max rel diff vs first variant %.1e)"
              % (inputs, batch, dt, v, pf, ms * 1e3, err), flush=True)
