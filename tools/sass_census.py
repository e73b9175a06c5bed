"""Per-kernel instruction census of libevt.so from `cuobjdump -sass` / `-res-usage` (no GPU needed) -> markdown.

    python tools/sass_census.py [edgevisiontransformer_b200/libevt.so] > profiles/r02_sass_census.md

Columns: tcgen05.mma (UTCHMMA, of which cta_group::2 = .2CTA), tcgen05.ld / st (LDTM / STTM), TMA loads / stores /
reduce-adds (UTMALDG / UTMASTG / UTMAREDG), legacy tensor-core MMAs (HMMA = mma.sync), MUFU, the ELECT + BRA.U.ANY
"waterfall" loops ptxas wraps around a uniform-datapath instruction whose operands it cannot prove warp-uniform
(0 for every MMA / TMA issue path since the roles are chosen with elect.sync), registers per thread and static shared memory.
"""
import collections
import os
import re
import subprocess
import sys

PATTERNS = [("UTCHMMA", r"\bUTC[A-Z]*MMA\b"), (".2CTA", r"\bUTC[A-Z]*MMA\.2CTA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
            ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UTMAREDG", r"\bUTMAREDG"), ("HMMA", r"\bHMMA"),
            ("MUFU", r"\bMUFU"), ("waterfall", r"BRA\.U\.ANY")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = []
    for n in out:
        n = re.sub(r"\(anonymous namespace\)::", "", n)
        n = re.sub(r"^void ", "", n)
        n = re.sub(r"\(.*$", "", n)
        n = n.replace("evt::", "")
        short.append(n)
    return short


def main(path):
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)))
            cur = None
    counts = collections.OrderedDict()
    cur = None
    arch = set()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch.add(m.group(1))
        if cur is None or "/*" not in line:
            continue
        counts[cur]["instr"] += 1 if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line) else 0
        for name, pat in PATTERNS:
            if re.search(pat, line):
                counts[cur][name] += 1
    names = list(counts)
    short = demangle(names)
    print(f"# SASS census of `{os.path.relpath(path)}` ({', '.join(sorted(arch)) or 'sm_100a'} cubins; `python tools/sass_census.py`)\n")
    print("Kernels with tensor-core, TMEM or TMA instructions (the remaining kernels are plain load / store / warp-reduction code).  "
          "`HMMA` = legacy `mma.sync`; `waterfall` = ELECT/BRA.U.ANY loops around uniform-datapath instructions.\n")
    hdr = ["kernel"] + [p[0] for p in PATTERNS] + ["SASS instr", "regs", "static smem"]
    print("| " + " | ".join(hdr) + " |")
    print("|" + "---|" * len(hdr))
    tot = collections.Counter()
    rows = []
    for n, s in zip(names, short):
        c = counts[n]
        tot.update(c)
        if not any(c[p[0]] for p in PATTERNS if p[0] != "MUFU"):
            continue
        r, sm = usage.get(n, (0, 0))
        rows.append((s, [c[p[0]] for p in PATTERNS] + [c["instr"], r, sm]))
    for s, vals in sorted(rows):
        print("| `" + s + "` | " + " | ".join(str(v) for v in vals) + " |")
    print("| **all %d kernels** | " % len(names) + " | ".join(str(tot[p[0]]) for p in PATTERNS) + f" | {tot['instr']} | | |")


if __name__ == "__main__":
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "edgevisiontransformer_b200", "libevt.so"))
