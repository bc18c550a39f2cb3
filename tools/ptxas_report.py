"""Compact per-kernel register / spill / shared-memory report from `nvcc -Xptxas -v`.
    python tools/ptxas_report.py [substring]
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util  # noqa: E402
_spec = importlib.util.spec_from_file_location("simpb_b200_build", os.path.join(ROOT, "simpb_b200", "build.py"))
build = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(build)

_, err = build.compile_objects(verbose=True, obj_dir="/tmp/dfa_ptxas_report")
filt = sys.argv[1] if len(sys.argv) > 1 else ""
name = None
for line in err.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("void ", "")
        name = re.sub(r"\(.*", "", name)
        spill = ""
    m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and (m.group(1) != "0" or m.group(2) != "0"):
        spill = " SPILL st=%s ld=%s" % m.groups()
    m = re.search(r"Used (\d+) registers", line)
    if m and name and filt in name:
        print("%-70s regs=%s%s" % (name, m.group(1), spill))
