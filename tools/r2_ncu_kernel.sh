#!/bin/bash
# ncu --set full of one kernel (regex $1) inside tools/profile_models.py $2 (after the same command ran clean)
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python tools/profile_models.py $2 > $O/k_plain.log 2>&1 || { tail -5 $O/k_plain.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$1" -c 2 -o $O/r02_kernel python tools/profile_models.py $2 > $O/k_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $O/k_ncu.log
