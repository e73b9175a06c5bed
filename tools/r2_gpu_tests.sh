#!/bin/bash
# full GPU test suite + short bench (development helper)
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > $O/gpu_tests.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests.log
tail -12 $O/gpu_tests.log
timeout 600 python -m edgevisiontransformer_b200.benchmark --model t2t_vit_14 --batch 256 --num_runs 10 --warmup_runs 5 | tail -1
