"""ncu target: two forwards each of the non-headline configs (Swin-T bs256, T2T-ViT-14 bs256, pruned DeiT-Tiny bs1024, DeiT-Small
bs256), separated by a `sign_` marker kernel so that `tools/launch_shares.py` can split the launch list per model.
No timing here -- numbers printed under a profiler are never bench values."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from edgevisiontransformer_b200 import B200ViTForImageClassification  # noqa: E402
from edgevisiontransformer_b200.benchmark.b200 import build_model  # noqa: E402

which = sys.argv[1:] or ["swin_tiny", "t2t_vit_14", "pruned_tiny", "deit_small"]
marker = torch.zeros(64, device="cuda")
for name in which:
    if name == "pruned_tiny":
        sd, heads, inter = bench.pruned_tiny_state_dict("h1_d230")
        m = B200ViTForImageClassification.from_state_dict(sd, device="cuda", max_batch=1024, keep_params=False)
        shape, batch = (3, 224, 224), 1024
    else:
        m, shape, _ = build_model(name, max_batch=256)
        batch = 256
    x = torch.randn(batch, *shape, device="cuda")
    with torch.no_grad():
        m(x)
        torch.cuda.synchronize()
        marker.sign_()
        m(x)
        marker.sign_()
    torch.cuda.synchronize()
    print("ok", name)
    del m, x
    torch.cuda.empty_cache()
