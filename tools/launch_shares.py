"""Split an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of tools/profile_models.py at the `sign_` marker kernels
and print, per model, the kernel shares of ONE forward as a markdown table.
    python tools/launch_shares.py gpurun_out/models_launches.csv swin_tiny t2t_vit_14 pruned_tiny deit_small > profiles/rNN_model_shares.md"""
import csv
import re
import sys
from collections import OrderedDict


def rows(path):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
        yield r["Kernel Name"], us


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*\)$", "", name)
    return name.replace("evt::", "")


def main():
    path, names = sys.argv[1], sys.argv[2:]
    segs, cur, marks = [], [], 0
    for k, us in rows(path):
        if "sign_kernel_cuda" in k:
            marks += 1
            if marks % 2 == 0:
                segs.append(cur)
            cur = []
            continue
        cur.append((short(k), us))
    print("# Kernel shares of one forward per model (ncu launch list, serialised launches at boost clocks: compare shares)\n")
    for name, seg in zip(names, segs):
        agg = OrderedDict()
        for k, us in seg:
            n, t = agg.get(k, (0, 0.0))
            agg[k] = (n + 1, t + us)
        tot = sum(t for _, t in agg.values())
        print(f"## {name}: {len(seg)} launches, {tot / 1e3:.3f} ms summed\n")
        print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f} % | {t / n:.1f} |")
        print()


if __name__ == "__main__":
    main()
