#!/bin/bash
# round-end validation: full GPU test suite, smoke, then the bench line exactly as the driver runs it (N = 1)
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > $O/final_tests.log 2>&1; echo "tests rc=$?" >> $O/final_tests.log
tail -4 $O/final_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/final_smoke.log 2>&1; tail -2 $O/final_smoke.log
timeout 1500 python bench.py > $O/final_n1.log 2> $O/final_n1.err; echo "bench rc=$?"
tail -c 6000 $O/final_n1.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/final_ref.log 2>&1; tail -c 600 $O/final_ref.log
