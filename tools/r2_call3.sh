#!/bin/bash
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > $O/c3_tests.log 2>&1; echo "tests rc=$?" >> $O/c3_tests.log
tail -15 $O/c3_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/c3_bench.log 2> $O/c3_bench.err; echo "bench rc=$?"; tail -c 6000 $O/c3_bench.log; tail -5 $O/c3_bench.err
