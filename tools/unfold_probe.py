"""Time the 7x7x3 soft split + LayerNorm of T2T's first stage alone (256 images) -- measurement aid."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from edgevisiontransformer_b200 import ops  # noqa: E402

x = torch.randn(256, 224, 224, 3, device="cuda")
g, b = torch.ones(147, device="cuda"), torch.zeros(147, device="cuda")
ms = bench.timed_steps(lambda: ops.unfold_ln_nhwc(x, 7, 4, 2, g, b, 1e-5, ld=152), 20, warmup=3)
x2 = torch.randn(256, 56, 56, 64, device="cuda")
g2, b2 = torch.ones(576, device="cuda"), torch.zeros(576, device="cuda")
ms2 = bench.timed_steps(lambda: ops.unfold_ln_nhwc(x2, 3, 2, 1, g2, b2, 1e-5, ld=576), 20, warmup=3)
print("unfold 7x7x3 + LN: %.1f us   unfold 3x3x64 + LN: %.1f us" % (ms * 1e3, ms2 * 1e3))
