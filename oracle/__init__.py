"""CPU oracle for SimPB's deformable feature aggregation path.

TEST INFRASTRUCTURE ONLY — imported by `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py`.  The product package
`simpb_b200` never imports this package (a test enforces that).

Two layers:

* `dfa_oracle.c` (built with gcc into `oracle/_build/libdfa_oracle.so`): the op itself —
  forward, backward and the integer side channel — restating
  /root/reference/projects/mmdet3d_plugin/ops/src/deformable_aggregation_cuda.cu.
* `module_ref.py`: the Python-level pieces around the op (feature-map flattening,
  key-point generation, camera projection, weight softmax, the grid_sample CPU path),
  restating /root/reference/projects/mmdet3d_plugin/{ops/__init__.py,models/blocks.py,
  models/detection3d/blocks.py} with plain torch CPU ops.

Pinning: the reference ships no tests or golden vectors for this path, so the
oracle is pinned against outputs of the reference itself: (1) `tests/golden/*.npz`, generated in
the build container by importing the unmodified reference module
(`tests/golden/make_golden.py`), and (2) on the GPU box, against the unmodified reference CUDA
op compiled by `oracle/build_ref.py` into `oracle/_ref/`.
"""
from .op_ref import (  # noqa: F401
    build_oracle,
    forward,
    backward,
    distinct_rows,
)
