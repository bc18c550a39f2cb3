"""bench.py contract checks that need no GPU: the reference arm (the reference's grid_sample CPU path
on the host cores) prints ONE JSON line with the agreed keys, and the product arm refuses to run
without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True,
                          text=True, cwd=ROOT, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["metric"] == "deformable_aggregation_forward_queries_per_sec" and d["unit"] == "queries/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["value"] > 0 and abs(d["value"] - 900 / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_product_arm_fails_loudly_without_a_gpu():
    if torch.cuda.is_available():
        return                       # on a GPU box the arm runs; the driver measures it there
    r = run_bench("--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")], "no bench line may be printed without a GPU"


def test_reference_arm_runs_the_references_own_code_when_staged():
    """With the reference's files present (/root/reference here, oracle/_ref/py on the GPU box) the arm's
    step is the reference's feature_sampling + multi_view_level_fusion; it agrees with the port the arm
    falls back to, and neither imports the product library."""
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    from oracle import ref_import
    if ref_import.default_root() is None:
        import pytest
        pytest.skip("reference files neither at /root/reference nor staged under oracle/_ref/py")
    a = argparse.Namespace(inputs="rig", batch=1, anchors=60, workload="fwd")
    step, A, kind = bench.cpu_reference_setup(a)
    assert kind == "reference" and A == 60
    out = step()
    saved = ref_import.default_root
    ref_import.default_root = lambda: None
    try:
        step2, _, kind2 = bench.cpu_reference_setup(a)
    finally:
        ref_import.default_root = saved
    assert kind2 == "port"
    out2 = step2()
    assert float((out - out2).abs().max() / out2.abs().max()) <= 5e-5


def test_reference_arm_does_not_load_the_product_library():
    code = ("import sys; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','0'];"
            "import runpy; runpy.run_path(%r, run_name='__main__');"
            "import os; maps=open('/proc/self/maps').read();"
            "assert 'libdfa_b200' not in maps, 'reference arm mapped the product library';"
            "assert 'simpb_b200.cabi' not in sys.modules") % os.path.join(ROOT, "bench.py")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
