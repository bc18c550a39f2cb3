"""Counts of the SASS mnemonics that prove a Blackwell-native build, per default kernel of
libdfa_b200.so (cuobjdump -sass; no GPU needed).   python tools/sass_extract.py > profiles/sass_r2.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "simpb_b200", "libdfa_b200.so")
WANT = ["UBLKCP", "SYNCS", "FFMA2", "LDG.E.128", "REDG.E.ADD.F32", "REDUX", "CREDUX", "VOTE", "MATCH", "BAR.SYNC",
        "LDS", "STS", "HMMA", "UTCHMMA"]
KERNELS = ["dfa_fwd_rows_kernelIfLi1ELb1ELi256ELi6", "dfa_fwd_win_kernelIfLi2ELb1ELi2ELi16",
           "dfa_fwd_fused_kernelIfLi256ELi6", "dfa_bwd_merge_kernelIfLi2ELi8ELi4ELi4ELb1ELi8",
           "msda_fwd_kernel", "dfa_flatten_level_vec_kernel"]
# REDG.E.ADD.F32 with a 128-bit operand (`.128`-less mnemonic, four registers) is red.global.add.v4.f32
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        for w in WANT:
            if op.startswith(w):
                counts[cur][w] += 1
                total[w] += 1
        counts[cur]["(instructions)"] += 1
print("libdfa_b200.so, sm_100a, cuobjdump -sass: mnemonic counts")
print("whole library:", dict(total))
for k in KERNELS:
    for name in counts:
        if k in name:
            print("%-60s %s" % (name[-60:], dict(counts[name])))
            break
