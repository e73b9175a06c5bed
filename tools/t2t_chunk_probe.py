"""T2T-ViT-14 throughput at batch 1024 for different chunk sizes (max_batch) -- measurement aid."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from edgevisiontransformer_b200.benchmark.b200 import _random_t2t_weights  # noqa: E402
from edgevisiontransformer_b200.modeling_t2t import B200T2TViT  # noqa: E402

x = torch.randn(1024, 224, 224, 3, device="cuda")
for chunk in (256, 512, 1024, 256):
    m = B200T2TViT(_random_t2t_weights(384, 14, 6, 3.0), depth=14, num_heads=6, device="cuda", max_batch=chunk)
    ms = bench.timed_steps(lambda: m(x).logits, 10, warmup=3)
    print("chunk", chunk, "img/s", round(1024 / ms * 1e3), "workspace MB", round(m.core._ws.numel() / 1e6) if getattr(m.core, "_ws", None) is not None else None, flush=True)
    del m
    torch.cuda.empty_cache()
