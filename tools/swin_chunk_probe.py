"""Swin-T throughput at batch 1024 for different chunk sizes (max_batch) -- measurement aid."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from transformers import SwinConfig, SwinForImageClassification  # noqa: E402
from edgevisiontransformer_b200.modeling_swin import B200SwinForImageClassification  # noqa: E402

torch.manual_seed(0)
hf = SwinForImageClassification(SwinConfig(image_size=224, patch_size=4, window_size=7, embed_dim=96, depths=[2, 2, 6, 2],
                                           num_heads=[3, 6, 12, 24], num_labels=1000)).eval()
x = torch.randn(1024, 3, 224, 224, device="cuda")
for chunk in (256, 512, 1024, 256):
    m = B200SwinForImageClassification.from_hf(hf, device="cuda", max_batch=chunk)
    ms = bench.timed_steps(lambda: m(x).logits, 10, warmup=3)
    print("chunk", chunk, "img/s", round(1024 / ms * 1e3), flush=True)
    del m
    torch.cuda.empty_cache()
