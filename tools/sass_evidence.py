"""SASS evidence for profiles/: per-kernel counts of the Blackwell-specific mnemonics (TMA: UTMALDG / UBLKCP; mbarrier:
SYNCS; fp64: DFMA; atomics: REDG; release/acquire accesses of the peer-memory protocol) in the built library.
Usage: python tools/sass_evidence.py > profiles/r2_sass_evidence.txt   (no GPU needed)"""
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ["UTMALDG", "UBLKCP", "SYNCS", "DFMA", "DMUL", "REDG", "ATOM", "SHFL", "LDS", "LDG", "STG", "ST.E", "LD.E",
             "MEMBAR", "NANOSLEEP", "BAR.SYNC", "BAR.RED"]


def main():
    for lib in ("libexab200.so", "libexahost.so"):
        path = os.path.join(ROOT, "exaconstit_b200", "lib", lib)
        out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
        print("== %s  (cuobjdump -sass; arch %s)" % (lib, ", ".join(arch)))
        cur, counts, size = None, defaultdict(lambda: defaultdict(int)), defaultdict(int)
        for line in out.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                cur = re.sub(r"\(.*", "", cur).replace("void ", "")
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if m and cur:
                size[cur] += 1
                for k in MNEMONICS:
                    if m.group(2).startswith(k):
                        counts[cur][k] += 1
        tot = defaultdict(int)
        for fn in sorted(size, key=lambda f: -size[f]):
            c = counts[fn]
            for k, v in c.items():
                tot[k] += v
            keys = [k for k in ("UTMALDG", "UBLKCP", "SYNCS", "REDG", "DFMA", "SHFL", "NANOSLEEP") if c.get(k)]
            print("  %-78s %6d instr  %s" % (fn[:78], size[fn], "  ".join("%s %d" % (k, c[k]) for k in keys)))
        print("  totals: " + "  ".join("%s %d" % (k, tot[k]) for k in MNEMONICS if tot.get(k)))
        print()


if __name__ == "__main__":
    main()
