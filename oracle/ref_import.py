"""Imports the UNMODIFIED reference Python of the hot path (test infrastructure only).

The reference plugin cannot be imported as a package without mmcv / mmdet (its __init__ star-imports
them), so the files on the path are imported with (a) empty package objects pre-seeded in sys.modules
and (b) a minimal stand-in for the handful of mmcv symbols they use (Linear = nn.Linear, registries,
init helpers).  No reference arithmetic is replaced.  Files imported, relative to `root`:
    projects/mmdet3d_plugin/models/blocks.py               (DeformableFeatureAggregation)
    projects/mmdet3d_plugin/models/detection3d/blocks.py   (SparseBox3DKeyPointsGenerator)
    projects/mmdet3d_plugin/core/box3d.py                  (index constants)
    projects/mmdet3d_plugin/ops/__init__.py                (feature_maps_format)
`root` is /root/reference in the build container (tests/golden/make_golden.py) or the copy that
oracle/build_ref.py stages under oracle/_ref/py (git-ignored; it travels to the GPU box so that
bench.py's reference arm times the reference's own code there).
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(HERE, "_ref", "py")
FILES = ("projects/mmdet3d_plugin/models/blocks.py", "projects/mmdet3d_plugin/models/detection3d/blocks.py",
         "projects/mmdet3d_plugin/core/box3d.py", "projects/mmdet3d_plugin/ops/__init__.py")


def available(root):
    return all(os.path.exists(os.path.join(root, f)) for f in FILES)


def default_root():
    """/root/reference where it exists, else the staged copy, else None."""
    for r in ("/root/reference", STAGED_ROOT):
        if available(r):
            return r
    return None


def import_reference(root):
    """Returns (models.blocks module, ops module) of the reference tree at `root`."""
    def pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [root + path]
        sys.modules[name] = m

    for n, p in [("projects", "/projects"), ("projects.mmdet3d_plugin", "/projects/mmdet3d_plugin"),
                 ("projects.mmdet3d_plugin.core", "/projects/mmdet3d_plugin/core"),
                 ("projects.mmdet3d_plugin.models", "/projects/mmdet3d_plugin/models"),
                 ("projects.mmdet3d_plugin.models.detection3d",
                  "/projects/mmdet3d_plugin/models/detection3d")]:
        pkg(n, p)

    class Registry:
        def __init__(self):
            self.d = {}

        def register_module(self, *a, **k):
            def deco(c):
                self.d[c.__name__] = c
                return c
            return deco

    def build_from_cfg(cfg, reg, default_args=None):
        cfg = dict(cfg)
        return reg.d[cfg.pop("type")](**cfg)

    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()

    class Scale(nn.Module):
        def __init__(self, scale=1.0):
            super().__init__()
            self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))

        def forward(self, x):
            return x * self.scale

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m

    regs = [Registry() for _ in range(4)]
    mod("mmcv")
    mod("mmcv.runner")
    mod("mmcv.cnn.bricks")
    mod("mmcv.cnn", Linear=nn.Linear, Scale=Scale,
        build_activation_layer=lambda c: nn.ReLU(inplace=True),
        build_norm_layer=lambda c, n: (None, nn.LayerNorm(n)),
        xavier_init=lambda m, distribution="normal", bias=0.0, gain=1: (
            nn.init.xavier_uniform_(m.weight, gain=gain), nn.init.constant_(m.bias, bias)),
        constant_init=lambda m, val, bias=0.0: (
            nn.init.constant_(m.weight, val), nn.init.constant_(m.bias, bias)),
        bias_init_with_prob=lambda p: float(-np.log((1 - p) / p)))
    mod("mmcv.runner.base_module", Sequential=nn.Sequential, BaseModule=BaseModule)
    mod("mmcv.cnn.bricks.transformer", FFN=object)
    mod("mmcv.utils", build_from_cfg=build_from_cfg)
    mod("mmcv.cnn.bricks.drop", build_dropout=lambda c: nn.Dropout(c.get("drop_prob", 0.0)))
    mod("mmcv.cnn.bricks.registry", ATTENTION=regs[0], PLUGIN_LAYERS=regs[1],
        FEEDFORWARD_NETWORK=regs[2], POSITIONAL_ENCODING=regs[3])
    blocks = importlib.import_module("projects.mmdet3d_plugin.models.blocks")
    importlib.import_module("projects.mmdet3d_plugin.models.detection3d.blocks")
    # feature_maps_format lives in ops/__init__.py, whose first line imports the CUDA
    # extension wrapper; load the function from the file with that import satisfied by a stub.
    sys.modules["projects.mmdet3d_plugin.ops.deformable_aggregation"] = types.SimpleNamespace(
        DeformableAggregationFunction=None)
    spec = importlib.util.spec_from_file_location(
        "projects.mmdet3d_plugin.ops", root + "/projects/mmdet3d_plugin/ops/__init__.py",
        submodule_search_locations=[])
    ops = importlib.util.module_from_spec(spec)
    sys.modules["projects.mmdet3d_plugin.ops"] = ops
    spec.loader.exec_module(ops)
    return blocks, ops
