#!/bin/bash
# plain run first, then the ncu launch list of the same command
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python tools/profile_models.py > $O/models_plain.log 2>&1 || { echo plain failed; tail -20 $O/models_plain.log; exit 1; }
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/models_launches.csv \
    python tools/profile_models.py > $O/models_ncu.log 2>&1
echo "ncu rc=$?"
tail -3 $O/models_ncu.log
