"""Shared/global wavefront hot spots per SASS instruction from an .ncu-rep source page.
    python tools/ncu_smem.py report.ncu-rep [top N]
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
col = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[col["Address"]] == "Address":
        if body:
            break
        continue
    body.append(r)


def f(r, name):
    try:
        return float(r[col[name]] or 0)
    except ValueError:
        return 0.0


tot_s = sum(f(r, "L1 Wavefronts Shared") for r in body)
tot_i = sum(f(r, "L1 Wavefronts Shared Ideal") for r in body)
tot_g = sum(f(r, "L2 Theoretical Sectors Global") for r in body)
tot_ex = sum(f(r, "Instructions Executed") for r in body)
print("shared wavefronts %d (ideal %d)   global sectors %d   warp instructions %d" % (tot_s, tot_i, tot_g, tot_ex))
order = sorted(range(len(body)), key=lambda i: -f(body[i], "L1 Wavefronts Shared"))[:top]
for i in sorted(order):
    r = body[i]
    print("%4d shared=%-9d ideal=%-9d exec=%-8d %s" % (i, f(r, "L1 Wavefronts Shared"), f(r, "L1 Wavefronts Shared Ideal"),
                                                      f(r, "Instructions Executed"), r[col["Source"]].strip()[:70]))
