#!/bin/bash
# same-box A/B: tanh-GELU epilogue with one MUFU per element (tanh.approx, libevt_tanh.so) against ex2 + rcp (shipped)
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/tanh_*.log
for lib in libevt libevt_tanh libevt libevt_tanh; do
  EVT_LIB_PATH=$PWD/edgevisiontransformer_b200/$lib.so timeout 300 python tools/config_sweep.py --steps 10 --only t2t >> $O/tanh_$lib.log 2>&1
done
for lib in libevt libevt_tanh; do
  EVT_LIB_PATH=$PWD/edgevisiontransformer_b200/$lib.so timeout 600 python -m pytest tests -q -m gpu -s -k "t2t or tf_dialect or tf_vit or gelu or ffn or FeedForward or layers" > $O/tanh_tests_$lib.log 2>&1; echo "rc=$?" >> $O/tanh_tests_$lib.log
  grep -h "max_abs\|passed\|failed\|rc=" $O/tanh_tests_$lib.log | tail -8
done
grep -h img_per_s $O/tanh_libevt.log $O/tanh_libevt_tanh.log | cut -c1-150
