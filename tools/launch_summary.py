"""Aggregates an ncu launch list (gpu__time_duration.sum per launch, --csv) per kernel: count, total, share."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    import gzip
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt", newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    iN, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if len(r) <= iV or r[iM] != "gpu__time_duration.sum":
            continue
        v = float(r[iV].replace(",", ""))
        u = r[iU]
        ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(u, 1.0)
        name = re.sub(r"\(.*", "", r[iN]).strip()
        name = re.sub(r"^void\s+", "", name)
        agg[name][0] += 1
        agg[name][1] += ns
    tot = sum(v[1] for v in agg.values())
    print("%-60s %8s %12s %8s %10s" % ("kernel", "launches", "total ms", "share", "avg us"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s %8d %12.3f %7.1f%% %10.2f" % (k[:60], n, t / 1e6, 100 * t / tot, t / n / 1e3))
    print("%-60s %8d %12.3f" % ("TOTAL", sum(v[0] for v in agg.values()), tot / 1e6))


if __name__ == "__main__":
    main(sys.argv[1])
