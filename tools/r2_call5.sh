#!/bin/bash
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "layernorm_fused or residual_layernorm" > $O/c5_ops.log 2>&1; echo "rc=$?" >> $O/c5_ops.log; tail -15 $O/c5_ops.log
timeout 1800 python -m pytest tests -q -m gpu > $O/c5_tests.log 2>&1; echo "tests rc=$?" >> $O/c5_tests.log
tail -12 $O/c5_tests.log
timeout 600 python tools/config_sweep.py --steps 10 --only small,pruned,t2t > $O/c5_sweep.log 2>&1; tail -5 $O/c5_sweep.log
EVT_FUSE_LN_A=0 timeout 600 python tools/config_sweep.py --steps 10 --only small,pruned,t2t > $O/c5_sweep_off.log 2>&1; tail -5 $O/c5_sweep_off.log
