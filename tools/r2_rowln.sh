#!/bin/bash
# development helper: the row-resident residual-GEMM + LayerNorm kernel -- op tests, model tests, then an on/off sweep
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "rows_in_tmem" > $O/rowln_ops.log 2>&1; echo "ops rc=$?" >> $O/rowln_ops.log
tail -15 $O/rowln_ops.log
grep -q "ops rc=0" $O/rowln_ops.log || exit 1
timeout 1500 python -m pytest tests -q -m gpu -x > $O/rowln_tests.log 2>&1; echo "tests rc=$?" >> $O/rowln_tests.log
tail -15 $O/rowln_tests.log
for v in 1 0; do
  EVT_FUSE_ROWLN=$v timeout 600 python tools/config_sweep.py --steps 10 --only small,pruned,t2t > $O/rowln_sweep_$v.log 2>&1
  tail -3 $O/rowln_sweep_$v.log | cut -c1-400
done
