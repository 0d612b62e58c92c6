"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, mean duration and share.

    python tools/launch_shares.py gpurun_out/launches.csv > profiles/rNN_launches.md
"""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rd:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki].replace("void ", "").replace("<unnamed>::", ""))
        v = float(r[vi].replace(",", ""))
        if r[ui] in ("ns", "nsecond"):
            v /= 1000.0
        elif r[ui] in ("ms", "msecond"):
            v *= 1000.0
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | mean us | total us | share |")
    print("|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k, n, t / n, t, 100 * t / tot))


if __name__ == "__main__":
    main(sys.argv[1])
