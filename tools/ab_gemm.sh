# A/B of differently built libevt.so files on the encoder GEMM shapes at M = 1024 * 197 (development aid).
#   usage: bash tools/ab_gemm.sh path/to/libevt_a.so ...      (the in-tree library runs last in every round)
run() { python - <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
from tools.gemm_bench import gemm
M = 1024 * 197
gemm(M, 3072, 768, act="gelu_erf", tag="fc1")
gemm(M, 2304, 768, tag="qkv")
gemm(M, 768, 3072, out_dtype=torch.float32, residual=True, tag="fc2")
PY
}
for i in 1 2; do
  for v in "$@"; do echo "== $v"; EVT_LIB_PATH=$v run 2>&1 | cut -c1-95; done
  echo "== in-tree"; run 2>&1 | cut -c1-95
done
