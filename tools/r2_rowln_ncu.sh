#!/bin/bash
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "rows_in_tmem" > $O/rowln_ops.log 2>&1; echo "ops rc=$?" >> $O/rowln_ops.log
tail -3 $O/rowln_ops.log
grep -q "ops rc=0" $O/rowln_ops.log || exit 1
timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_model.py -q -x > $O/rowln_tests.log 2>&1; echo "tests rc=$?" >> $O/rowln_tests.log
tail -4 $O/rowln_tests.log
EVT_FUSE_ROWLN=1 timeout 600 python tools/config_sweep.py --steps 10 --only small,pruned,t2t > $O/rowln_sweep_1.log 2>&1
grep -c "timed out" $O/rowln_sweep_1.log
tail -3 $O/rowln_sweep_1.log | cut -c1-200
timeout 300 python tools/profile_rowln.py > $O/rowln_plain.log 2>&1 || { tail -5 $O/rowln_plain.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_rowln|gemm_pair|ln_rows" -o $O/r02_rowln python tools/profile_rowln.py 1 > $O/rowln_ncu.log 2>&1
echo "ncu rc=$?"
tail -3 $O/rowln_ncu.log
