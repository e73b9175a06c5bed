#!/bin/bash
# ncu --set full over one forward of a model of tools/profile_models.py ($1 = model, $2 = launches to skip, $3 = launches to take)
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python tools/profile_models.py $1 > $O/m_plain.log 2>&1 || { tail -5 $O/m_plain.log; exit 1; }
timeout 1500 ncu --set full --clock-control none --launch-skip $2 --launch-count $3 -o $O/r02_model_$1 python tools/profile_models.py $1 > $O/m_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $O/m_ncu.log
