#!/bin/bash
set -x
cd "$(dirname "$0")/.."
for lib in libevt libevt_u2 libevt_u4 libevt libevt_u2 libevt_u4; do
  EVT_LIB_PATH=$PWD/edgevisiontransformer_b200/$lib.so timeout 600 python -m edgevisiontransformer_b200.benchmark --model t2t_vit_14 --batch 256 --num_runs 30 --warmup_runs 5 2>&1 | tail -1 | cut -c1-200
done
