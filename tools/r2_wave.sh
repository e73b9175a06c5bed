#!/bin/bash
# same-box A/B: wave-aware tile width (EVT_GEMM_WAVE, csrc/gemm.cu choose_bn) on DeiT-Small; chunk size of the DeiT-Base step
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/wave_*.log
for w in 0 1 0 1; do
  EVT_GEMM_WAVE=$w timeout 300 python tools/config_sweep.py --steps 20 --only small >> $O/wave_$w.log 2>&1
done
for c in 1024 2048 1024 2048; do
  timeout 300 python bench.py --chunk $c --no-extras --steps 5 --warmup 3 >> $O/wave_chunk_$c.log 2>&1
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/wave_*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            if "config" in d and "img_per_s" in d:
                print(f, d["config"], round(d["img_per_s"]), {k: v["us_per_launch"] for k, v in d.get("stages", {}).items()})
            else:
                print(f, round(d["value"]), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
PY
