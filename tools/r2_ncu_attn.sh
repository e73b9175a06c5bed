#!/bin/bash
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 120 python tools/attn_probe.py 1024 197 12 > $O/ncu_attn2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention2_kernel -s 3 -c 1 -o $O/attn2_prof -f python tools/attn_probe.py 1024 197 12 > $O/ncu_attn2.log 2>&1
tail -3 $O/ncu_attn2.log
