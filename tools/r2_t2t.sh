#!/bin/bash
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_guards.py tests/test_gpu_model.py tests/test_gpu_full_size.py -q -x -k "unfold or t2t or performer" > $O/t2t_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/t2t_tests.log
grep -n "^E " $O/t2t_tests.log | head
for i in 1 2; do
  timeout 600 python -m edgevisiontransformer_b200.benchmark --model t2t_vit_14 --batch 1024 --num_runs 10 --warmup_runs 3 2>&1 | tail -1 | cut -c60-200
done
