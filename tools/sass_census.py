#!/usr/bin/env python
"""Static SASS instruction census of every kernel in the built objects (no GPU needed):
   python tools/sass_census.py > profiles/rNN_sass_census.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["VABSDIFF4", "LDG.E.128", "LDG", "STG.E.128", "STG", "LDS", "STS", "ATOMS", "ATOMG", "ATOM", "RED", "REDUX", "MATCH", "VOTE", "SHFL",
        "BAR", "DFMA", "DMUL", "DADD", "MUFU", "IMAD", "PRMT", "VIMNMX", "HMMA", "UTCMMA", "UTMALDG"]
EXCLUDE = {"LDG": "LDG.E.128", "STG": "STG.E.128", "RED": "REDUX"}

print("# SASS census of the sm_100a kernels (cuobjdump -sass of the built objects): static instruction counts per kernel;")
print("# the SAD kernels are VABSDIFF4.U8.ACC chains, vector loads are 128-bit, warp collectives are REDUX / MATCH / VOTE / SHFL;")
print("# no tensor-core instruction appears anywhere (none of the stages is a dense contraction, DESIGN.md 5); the descriptor kernel stages its tile with one TMA load (UTMALDG).")
for obj in sorted(glob.glob(os.path.join(ROOT, "low-cost*", "lib", "obj", "k_*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn, cnt = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            cnt[fn] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            cnt[fn][m.group(2)] += 1
            cnt[fn]["_total"] += 1
    for fn, c in cnt.items():
        name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", re.sub(r"svb::\(anonymous namespace\)::", "", name))
        parts = []
        for k in KEYS:
            n = 0
            for op, v in c.items():
                if not op.startswith(k) or (k in EXCLUDE and op.startswith(EXCLUDE[k])):
                    continue
                if k == "ATOM" and (op.startswith("ATOMS") or op.startswith("ATOMG")):
                    continue
                n += v
            if n:
                parts.append("%s=%d" % (k, n))
        print("%-34s %5d instr  %s" % (name[:34], c["_total"], " ".join(parts)))
