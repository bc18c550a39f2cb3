"""Hottest SASS instructions of an .ncu-rep by warp-stall samples (needs --import-source on).
    python tools/ncu_hot.py report.ncu-rep [top N]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[ia] == "Address":   # a second kernel's block starts: first one only
        if body:
            break
        continue
    body.append(r)
total = sum(int(r[ismp] or 0) for r in body)
print("total samples", total, " instructions", len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ismp] or 0))[:top]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
    print("%4d %5.1f%% exec=%-7s %-60s %s" % (i, 100.0 * int(r[ismp] or 0) / max(total, 1), r[iex],
                                             r[isrc].strip()[:60], " ".join("%s=%d" % (n, v) for v, n in st if v)))
