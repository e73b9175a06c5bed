#!/bin/bash
# ncu --set full of the row-resident GEMM + LayerNorm kernel and of the two kernels it replaces (after the same command ran clean)
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python tools/profile_rowln.py > $O/rowln_plain.log 2>&1 || { tail -5 $O/rowln_plain.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_rowln|gemm_pair|ln_rows" -o $O/r02_rowln python tools/profile_rowln.py 1 > $O/rowln_ncu.log 2>&1
echo "ncu rc=$?"
tail -3 $O/rowln_ncu.log
