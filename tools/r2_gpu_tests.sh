#!/bin/bash
# full GPU test suite + Swin / T2T quick benchmarks (development helper)
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > $O/gpu_tests.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests.log
tail -12 $O/gpu_tests.log
timeout 600 python -m edgevisiontransformer_b200.benchmark --model swin_tiny --batch 256 --num_runs 10 --warmup_runs 5 | tail -1
timeout 600 python -m edgevisiontransformer_b200.benchmark --model swin_tiny --batch 1 --graph --num_runs 50 --warmup_runs 20 | tail -1
