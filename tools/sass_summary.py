#!/usr/bin/env python
"""Counts the Blackwell-native SASS mnemonics per kernel of the in-tree library (B200_PROFILING.md "what proves a
Blackwell-native kernel"): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA loads /
stores / reduce-stores, UTCBAR = tcgen05.commit, SYNCS = mbarrier.     usage: python tools/sass_summary.py > profiles/sass_tc.txt"""
import os
import re
import subprocess
import sys
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mliis_b200", "libmliis_b200.so")
PAT = re.compile(r"\b(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UTCBAR|UTMACCTL|UBLKCP|HMMA|SYNCS)\b")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per, cur = defaultdict(Counter), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur).replace("mliis::", "")
            continue
        if cur:
            for t in PAT.findall(line):
                per[cur][t] += 1
    print("# SASS mnemonic counts per kernel of mliis_b200/libmliis_b200.so (cuobjdump -sass, sm_100a)")
    cols = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "HMMA"]
    print("%-40s " % "kernel" + " ".join("%9s" % c for c in cols))
    for k in sorted(per, key=lambda k: -sum(per[k].values())):
        c = per[k]
        mma = sum(v for n, v in c.items() if n.startswith("UTC") and n.endswith("MMA"))
        row = [mma, c["LDTM"], c["UTMALDG"], c["UTMASTG"], c["UTMAREDG"], c["UTCBAR"], c["SYNCS"], c["HMMA"]]
        if sum(row):
            print("%-40s " % k[:40] + " ".join("%9d" % v for v in row))


if __name__ == "__main__":
    main()
