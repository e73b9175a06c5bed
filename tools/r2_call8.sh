#!/bin/bash
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/c8_ops.log 2>&1; tail -2 $O/c8_ops.log
for rep in 1 2; do
  echo "== new"; timeout 300 python tools/gemm_bench.py base 1024
  echo "== prev"; EVT_LIB_PATH=$PWD/edgevisiontransformer_b200/libevt_prev.so timeout 300 python tools/gemm_bench.py base 1024
done > $O/c8_gemm.log 2>&1
grep -E "==|gemm|attention|layer total" $O/c8_gemm.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/c8_bench.log 2> $O/c8_bench.err
EVT_LIB_PATH=$PWD/edgevisiontransformer_b200/libevt_prev.so timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/c8_bench_prev.log 2> $O/c8_bench_prev.err
python - <<'PY'
import json
for f in ("gpurun_out/c8_bench.log","gpurun_out/c8_bench_prev.log"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, round(d["value"]), d["clocks"]["sm_mhz"], {k:round(v["ms_per_launch"],4) for k,v in d["stages"].items() if isinstance(v,dict)})
PY
