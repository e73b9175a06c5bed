#!/bin/bash
# development helper: full GPU tests, sweep of the narrow configs, then a same-box A/B of the headline bench (new vs previous library)
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > $O/ab_tests.log 2>&1; echo "tests rc=$?" >> $O/ab_tests.log
tail -5 $O/ab_tests.log
timeout 600 python tools/config_sweep.py --steps 10 --only small,pruned,t2t > $O/ab_sweep.log 2>&1
grep -c "timed out" $O/ab_sweep.log
cut -c1-170 $O/ab_sweep.log | tail -3
for lib in new prev new prev; do
  if [ $lib = prev ]; then export EVT_LIB_PATH=$PWD/edgevisiontransformer_b200/libevt_prev.so; else unset EVT_LIB_PATH; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2> $O/ab_bench_$lib.err | tail -1 > $O/ab_bench_$lib.log
  python - <<PY
import json
d=json.loads(open("$O/ab_bench_$lib.log").read())
print("$lib", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v.get("ms_per_launch", 0), 4) if isinstance(v, dict) else v for k, v in d.get("stages", {}).items()})
PY
done
