#!/usr/bin/env python
"""Opcode histogram of every shipped sm_100a object (cuobjdump -sass on csrc/build/*.o) -> profiles/r2_sass_opcodes.txt.
Runs without a GPU:  python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TENSOR = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCATOMSWS", "UTMACMDFLUSH", "UTMACCTL", "SYNCS", "ATOMS",
          "ATOMG", "RED", "REDG", "LDGSTS", "UBLKCP", "ELECT", "FENCE", "CCTL")


def main():
    objs = sorted(glob.glob(os.path.join(ROOT, "explorative-attention-vit-*_b200", "csrc", "build", "*.o")))
    assert objs, "build first: python -c 'import __graft_entry__ as g; g.build()'"
    print("# SASS opcode evidence per object (cuobjdump -sass on the shipped sm_100a objects; tools/sass_opcodes.py)")
    print("# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit, "
          "SYNCS = mbarrier, ATOMS/RED = atomics")
    print("# LD / ST without a space suffix (generic addressing) must not appear in the tcgen05 kernels: the shared-memory accesses "
          "are LDS / STS (DESIGN 3b.4)")
    for o in objs:
        sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True, check=True).stdout
        hist = collections.Counter()
        for line in sass.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?[A-Z]+\d*\s+)?([A-Z][A-Z0-9_]*)", line)
            if m:
                hist[m.group(1)] += 1
        print(f"\n## {os.path.basename(o)}")
        print("  ".join(f"{k}:{v}" for k, v in hist.most_common()))
        print("tensor-path opcodes: " + "  ".join(f"{k}:{hist[k]}" for k in TENSOR if hist.get(k)))
        print(f"generic LD:{hist.get('LD', 0)}  generic ST:{hist.get('ST', 0)}  LDS:{hist.get('LDS', 0)}  STS:{hist.get('STS', 0)}")


if __name__ == "__main__":
    sys.exit(main())
