"""Per-kernel budget of ONE proof from an `ncu --metrics gpu__time_duration.sum --csv` launch list of tools/l2_one_proof.py:
launch count, summed duration, and blocks x duration (what a small proof costs the GPU when many run side by side).

    python tools/launch_budget.py gpurun_out/l2_launches.csv > profiles/...txt
"""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
    names = [r[ki] for r in data]
    tails = [i for i, n in enumerate(names) if "prove_tail_scalars" in n]
    start = tails[-2]                      # the last proof starts with its two prove_tail_scalars launches
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in data[start:]:
        g = [int(x) for x in re.findall(r"\d+", r[gi])]
        nb = g[0] * g[1] * g[2]
        d = float(r[vi].replace(",", "")) / 1000.0
        n = re.sub(r"\(.*", "", r[ki])
        n = re.sub(r"void |zkb::|<unnamed>::|cub::|policy_hub<.*", "", n)[:52]
        a = agg[n]
        a[0] += 1
        a[1] += d
        a[2] += nb * d / 1000.0
    print("# last proof of %s: %d launches" % (path, len(data) - start))
    print("%-52s %5s %12s %14s" % ("kernel", "x", "sum us", "blocks x ms"))
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        print("%-52s %5d %12.1f %14.2f" % (n, a[0], a[1], a[2]))
    print("%-52s %5d %12.1f %14.2f" % ("TOTAL", sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values()),
                                       sum(a[2] for a in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1])
