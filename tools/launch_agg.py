"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, summed duration, share.
    python tools/launch_agg.py launches.csv"""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hdr], [r for r in rows[hdr + 1:] if len(r) > 5]
    ki, vi, gi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0, 0])
    order = []
    for r in data:
        d = float(r[vi].replace(",", ""))
        u = r[ui]
        d = d / 1e3 if u in ("ns", "nsecond") else (d if u in ("us", "usecond") else d * 1e3)
        n = re.sub(r"\(.*", "", r[ki])
        n = re.sub(r"void |zkb::|<unnamed>::|\(anonymous namespace\)::", "", n)[:60]
        g = [int(x) for x in re.findall(r"\d+", r[gi])]
        if n not in agg:
            order.append(n)
        a = agg[n]
        a[0] += 1
        a[1] += d
        a[2] = max(a[2], g[0] * g[1] * g[2])
    tot = sum(a[1] for a in agg.values())
    print("# %s: %d launches, %.1f us summed" % (path, len(data), tot))
    print("%-60s %5s %12s %7s %10s" % ("kernel", "x", "sum us", "share", "max grid"))
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s %5d %12.1f %6.1f%% %10d" % (n, a[0], a[1], 100 * a[1] / tot, a[2]))


if __name__ == "__main__":
    main(sys.argv[1])
