#!/bin/bash
# one-pass ncu metrics (duration + issue-slot utilisation + DRAM throughput) for every launch of tools/profile_models.py [models...]
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python tools/profile_models.py "$@" > $O/models_plain.log 2>&1 || { echo plain failed; tail -20 $O/models_plain.log; exit 1; }
timeout 1500 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file $O/models_issue.csv \
    python tools/profile_models.py "$@" > $O/models_ncu.log 2>&1
echo "ncu rc=$?"
ls -la $O/models_issue.csv
