#!/bin/bash
# development helper: epilogue-warp count of the row-resident GEMM + LayerNorm kernel (8 / 12 / 16)
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
for w in 8; do
  EVT_ROWLN_WARPS=$w EVT_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "rows_in_tmem" > $O/rowln_ops_$w.log 2>&1; echo "ops w=$w rc=$?"
  grep "evt: gemm_rowln" $O/rowln_ops_$w.log | sort -u
  tail -2 $O/rowln_ops_$w.log
  EVT_ROWLN_WARPS=$w timeout 600 python tools/config_sweep.py --steps 10 --only small,pruned,t2t > $O/rowln_sweep_w$w.log 2>&1
  grep -c "timed out" $O/rowln_sweep_w$w.log
  python - <<PY
import json
for l in open("$O/rowln_sweep_w$w.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("w=$w", d["config"], round(d["img_per_s"]), {k: v["us_per_launch"] for k, v in d.get("stages", {}).items() if k in ("out_proj", "fc2")})
PY
done
