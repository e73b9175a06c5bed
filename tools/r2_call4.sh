#!/bin/bash
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k attention > $O/c4_attn_tests.log 2>&1; echo "rc=$?" >> $O/c4_attn_tests.log; tail -3 $O/c4_attn_tests.log
for s in "1024 197 12" "256 197 6" "1024 197 3" "1024 198 12"; do
  echo "== rot $s"; timeout 120 python tools/attn_probe.py $s
  echo "== norot $s"; EVT_ATTN_NOROT=1 timeout 120 python tools/attn_probe.py $s
done > $O/c4_attn.log 2>&1
grep -v "^+" $O/c4_attn.log
timeout 1800 python -m pytest tests -q -m gpu > $O/c4_tests.log 2>&1; echo "tests rc=$?" >> $O/c4_tests.log
tail -12 $O/c4_tests.log
