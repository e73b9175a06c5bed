"""Library baseline on the same B200: the reference's own forward (HF ViTForImageClassification) in bf16 on the GPU
through stock PyTorch kernels (cuBLAS + SDPA).  Reported in DESIGN.md next to bench.py's number; not part of the product."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_hf

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
for impl in ("sdpa", "eager"):
    from transformers import ViTConfig, ViTForImageClassification
    cfg = ViTConfig(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, num_labels=1000,
                    image_size=224, patch_size=16, attn_implementation=impl)
    torch.manual_seed(0)
    model = ViTForImageClassification(cfg).eval().cuda().bfloat16()
    x = torch.randn(B, 3, 224, 224, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        for _ in range(3):
            model(pixel_values=x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 8
        e0.record()
        for _ in range(n):
            model(pixel_values=x).logits
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"torch {torch.__version__} HF ViT-Base bf16 attn={impl} batch {B}: {ms:.2f} ms/forward -> {B / ms * 1e3:.0f} img/s", flush=True)
    del model
