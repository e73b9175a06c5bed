#!/bin/bash
# Round-2 call 1: correctness of the elect.sync roles + attention2, then same-box A/B against the round-1 library.
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/ab1_smi.txt
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/ab1_ops.log 2>&1; echo "ops rc=$?" >> $O/ab1_ops.log
tail -5 $O/ab1_ops.log
for which in base small tiny; do
  B=1024; [ $which = small ] && B=256
  echo "== new $which" ; timeout 300 python tools/gemm_bench.py $which $B
  echo "== r1 $which" ; EVT_LIB_PATH=$PWD/edgevisiontransformer_b200/libevt_r1.so timeout 300 python tools/gemm_bench.py $which $B
done > $O/ab1_gemm.log 2>&1
for s in "1024 197 12" "256 197 6" "1024 197 1" "1024 198 12"; do
  echo "== v2 $s"; timeout 120 python tools/attn_probe.py $s
  echo "== v1 $s"; EVT_ATTN_V1=1 timeout 120 python tools/attn_probe.py $s
done > $O/ab1_attn.log 2>&1
cat $O/ab1_gemm.log $O/ab1_attn.log | grep -v "^+" | tail -80
timeout 1500 python -m pytest tests -x -q -m gpu > $O/ab1_tests.log 2>&1; echo "tests rc=$?" >> $O/ab1_tests.log
tail -5 $O/ab1_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/ab1_bench.log 2> $O/ab1_bench.err; tail -2 $O/ab1_bench.log
EVT_LIB_PATH=$PWD/edgevisiontransformer_b200/libevt_r1.so timeout 600 python bench.py --steps 10 --warmup 3 > $O/ab1_bench_r1.log 2> $O/ab1_bench_r1.err; tail -2 $O/ab1_bench_r1.log
