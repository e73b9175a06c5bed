#!/bin/bash
# Round-2 evidence: launch list of a bench step and ncu --set full of one encoder layer (after each plain run exits 0).
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python bench.py --global-batch 1024 --steps 2 --warmup 3 --no-extras --no-cpu-baseline --preroll-s 0.2 > $O/r02_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file $O/r02_launches_bs1024.csv \
    python bench.py --global-batch 1024 --steps 2 --warmup 3 --no-extras --no-cpu-baseline --preroll-s 0.2 > $O/r02_ncu_bench.log 2>&1
timeout 300 python tools/profile_target.py 1024 > $O/r02_plain_pt.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_|attention2_kernel|ln_rows" -s 14 -c 7 -f -o $O/r02_layer \
    python tools/profile_target.py 1024 > $O/r02_ncu_pt.log 2>&1
tail -3 $O/r02_ncu_pt.log
