"""Split an ncu launch list of tools/profile_models.py (`--metrics gpu__time_duration.sum[,smsp__issue_active...,dram__throughput...]
--csv`) at the `sign_` marker kernels and print, per model, the kernel shares of ONE forward as a markdown table (with the
time-weighted issue-slot and DRAM utilisation per kernel when the list carries them).
    python tools/launch_shares.py gpurun_out/models_issue.csv swin_tiny t2t_vit_14 pruned_tiny deit_small > profiles/rNN_model_shares.md"""
import collections
import csv
import re
import sys

ISSUE = "smsp__issue_active.avg.pct_of_peak_sustained_active"
DRAM = "dram__throughput.avg.pct_of_peak_sustained_elapsed"


def launches(path):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    out = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = out.setdefault(int(r["ID"]), {"k": r["Kernel Name"]})
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Name"] == "gpu__time_duration.sum":
            unit = r.get("Metric Unit", "ns")
            v = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
        d[r["Metric Name"]] = v
    return list(out.values())


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*\)$", "", name)
    return name.replace("evt::", "").replace("<unnamed>::", "")


def main():
    path, names = sys.argv[1], sys.argv[2:]
    segs, cur, marks = [], [], 0
    for d in launches(path):
        if "sign_kernel_cuda" in d["k"]:
            marks += 1
            if marks % 2 == 0:
                segs.append(cur)
            cur = []
            continue
        cur.append(d)
    extra = any(ISSUE in d for s in segs for d in s)
    print("# Kernel shares of one forward per model (ncu launch list; serialised launches at boost clocks: compare shares)\n")
    print("`tools/r2_models_issue.sh` (after the same command ran clean): `ncu --metrics gpu__time_duration.sum,smsp__issue_active...,")
    print("dram__throughput... --clock-control none --csv python tools/profile_models.py` = Swin-T batch 256, T2T-ViT-14 batch 256, pruned")
    print("DeiT-Tiny batch 1024, DeiT-Small batch 256.  `issue` = issue slots busy, `dram` = DRAM throughput, both time-weighted over the")
    print("kernel's launches: a kernel with issue >= 65 % is bound by its instruction count, one with dram >= 60 % by HBM (ncu flushes the")
    print("caches before each launch, so write-backs land outside the window and `dram` under-counts stores).\n")
    for name, seg in zip(names, segs):
        agg = collections.OrderedDict()
        for d in seg:
            a = agg.setdefault(short(d["k"]), [0, 0.0, 0.0, 0.0])
            t = d["gpu__time_duration.sum"]
            a[0] += 1
            a[1] += t
            a[2] += t * d.get(ISSUE, 0.0)
            a[3] += t * d.get(DRAM, 0.0)
        tot = sum(a[1] for a in agg.values())
        print(f"## {name}: {len(seg)} launches, {tot / 1e3:.3f} ms summed\n")
        print("| kernel | launches | total us | share | avg us |" + (" issue % | dram % |" if extra else ""))
        print("|---|---|---|---|---|" + ("---|---|" if extra else ""))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            row = f"| `{k}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f} % | {a[1] / a[0]:.1f} |"
            if extra:
                row += f" {a[2] / a[1]:.0f} | {a[3] / a[1]:.0f} |"
            print(row)
        print()


if __name__ == "__main__":
    main()
