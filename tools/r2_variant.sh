#!/bin/bash
# development helper: same-box A/B of a differently built library (EVT_BUILD_OUT / EVT_NVCC_EXTRA in build.py) on the narrow configs
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
V=${1:-libevt_stg2}
for lib in libevt $V libevt $V; do
  EVT_LIB_PATH=$PWD/edgevisiontransformer_b200/$lib.so timeout 600 python tools/config_sweep.py --steps 10 --only small,pruned,t2t,base > $O/variant_$lib.log 2>&1
  python - <<PY
import json
for l in open("$O/variant_$lib.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$lib", d["config"], round(d["img_per_s"]), {k: v["us_per_launch"] for k, v in d.get("stages", {}).items() if k in ("qkv", "out_proj", "fc1", "fc2")})
PY
done
