#!/bin/bash
# same-box A/B: window attention with the next item's K / V / Q requested under the current item's math (EVT_SWIN_ATTN_PREFETCH)
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/swinpf_*.log
timeout 900 python -m pytest tests/test_gpu_swin.py tests/test_gpu_guards.py -q -m gpu -x > $O/swinpf_tests.log 2>&1; echo "tests rc=$?" >> $O/swinpf_tests.log
tail -3 $O/swinpf_tests.log
for w in 0 1 0 1; do
  EVT_SWIN_ATTN_PREFETCH=$w timeout 300 python tools/config_sweep.py --steps 10 --only swin >> $O/swinpf_$w.log 2>&1
done
grep -h img_per_s $O/swinpf_0.log $O/swinpf_1.log | cut -c1-200
