#!/bin/bash
# same-box A/B: row-resident GEMM + LayerNorm kernel at D = 384 with four operand stages (libevt_s4.so) against the shipped three
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
for lib in libevt libevt_s4 libevt libevt_s4; do
  EVT_LIB_PATH=$PWD/edgevisiontransformer_b200/$lib.so timeout 300 python tools/config_sweep.py --steps 10 --only small,t2t >> $O/s4_$lib.log 2>&1
done
for lib in libevt libevt_s4; do
  EVT_ROWLN_ALWAYS=1 EVT_LIB_PATH=$PWD/edgevisiontransformer_b200/$lib.so timeout 300 python tools/config_sweep.py --steps 10 --only small >> $O/s4_always_$lib.log 2>&1
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/s4_*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d["config"], round(d["img_per_s"]), {k: v["us_per_launch"] for k, v in d.get("stages", {}).items() if k in ("qkv", "out_proj", "fc1", "fc2", "layernorm")})
PY
