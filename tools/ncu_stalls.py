"""Per-instruction stall summary of one kernel from `ncu -i X.ncu-rep --page source --csv` output (first kernel)."""
import csv
import sys
from collections import Counter


def main(path, ntop=40):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ia, isrc, isamp, iexec = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    data = []
    for r in rows[2:]:
        if len(r) <= isamp or r[0] == 'Kernel Name' or r[isamp] == '# Samples':
            if data:
                break
            continue
        data.append(r)
    tot = sum(int(r[isamp] or 0) for r in data)
    print("total samples", tot, "instrs", len(data), "warp-instr executed", sum(int(r[iexec] or 0) for r in data))
    agg = Counter()
    for r in data:
        for i in stall:
            agg[hdr[i]] += int(r[i] or 0)
    print(agg.most_common(12))
    for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:ntop]:
        st = {hdr[i][6:]: r[i] for i in stall if r[i] not in ('0', '')}
        print(r[ia][-5:], r[isamp].rjust(5), r[iexec].rjust(9), r[isrc][:70].ljust(70), st)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
