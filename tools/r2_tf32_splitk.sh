#!/bin/bash
# same-box A/B: split K for the residual GEMMs of the tf32 (accuracy) mode at batch 1 (EVT_TF32_SPLIT_K)
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/tf32sk_*.log
for w in 0 1 0 1; do
  EVT_TF32_SPLIT_K=$w timeout 300 python tools/config_sweep.py --only tiny_bs1 >> $O/tf32sk_$w.log 2>&1
done
EVT_TF32_SPLIT_K=1 timeout 600 python -m pytest tests -q -m gpu -x -k "tf32" > $O/tf32sk_tests.log 2>&1; echo "rc=$?" >> $O/tf32sk_tests.log
tail -3 $O/tf32sk_tests.log
grep -h ms_per_step $O/tf32sk_0.log $O/tf32sk_1.log | cut -c1-160
