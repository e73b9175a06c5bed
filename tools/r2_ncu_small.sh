#!/bin/bash
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python tools/profile_target.py 256 384 6 1536 > $O/small_plain_pt.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm_|attention2_kernel|ln_rows" -s 14 -c 7 -f -o $O/r02_small_layer \
    python tools/profile_target.py 256 384 6 1536 > $O/small_ncu_pt.log 2>&1
tail -3 $O/small_ncu_pt.log
