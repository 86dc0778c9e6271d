#!/usr/bin/env python
"""Fingerprint of the device code of every CUDA translation unit: md5 over the SASS instruction lines of lib/obj/*.o (addresses and
mnemonics, not the mangled names -- the anonymous-namespace hash in those depends on the source path).  Host-only changes leave it
untouched, so a fingerprint committed next to a round's GPU evidence says whether the kernels in the tree are the ones that were
measured:  python tools/sass_fingerprint.py [--check profiles/<file>]"""
import glob
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = glob.glob(os.path.join(ROOT, "low-cost-*_b200", "lib", "obj"))[0]
LINE = re.compile(r"^\s+/\*[0-9a-f]{4}\*/")


def fingerprints():
    out = {}
    for path in sorted(glob.glob(os.path.join(OBJ, "k_*.o")) + glob.glob(os.path.join(OBJ, "band_split.o")) + glob.glob(os.path.join(OBJ, "pipeline.o"))):
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
        lines = [l for l in sass.splitlines() if LINE.match(l)]
        out[os.path.basename(path)] = (hashlib.md5("\n".join(lines).encode()).hexdigest(), len(lines))
    return out


def main():
    fp = fingerprints()
    text = "".join("%-18s %s %7d instructions\n" % (k, h, n) for k, (h, n) in fp.items())
    if len(sys.argv) == 3 and sys.argv[1] == "--check":
        want = open(sys.argv[2]).read()
        want = "".join(l + "\n" for l in want.splitlines() if not l.startswith("#"))
        if want != text:
            sys.stdout.write(text)
            sys.exit("device code differs from " + sys.argv[2])
        print("device code equals", sys.argv[2])
        return
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
