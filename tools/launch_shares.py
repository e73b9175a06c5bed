"""Kernel shares from an `ncu --metrics gpu__time_duration.sum --csv --log-file X.csv` launch list (no GPU needed).

    python tools/launch_shares.py gpurun_out/launches.csv [--md]
"""
import collections
import csv
import re
import sys


def shares(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    order = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("unnamed>::", "").replace("evt::<", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        order.append((k, v))
    return agg, order


if __name__ == "__main__":
    agg, order = shares(sys.argv[1])
    tot = sum(a[1] for a in agg.values())
    md = "--md" in sys.argv
    print(f"{len(order)} launches, {tot:.1f} us")
    if md:
        print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if md:
            print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[1] / tot * 100:.1f} % | {a[1] / a[0]:.1f} |")
        else:
            print(f"{k[:72]:72s} {a[0]:4d} {a[1]:10.1f} {a[1] / tot * 100:5.1f}% {a[1] / a[0]:8.1f}")
    if "--seq" in sys.argv:
        for k, v in order:
            print(f"{v:9.1f}  {k[:90]}")
