"""Model shape specification shared by the oracle restatements (test infrastructure)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List


@dataclass
class ViTSpec:
    """Per-layer shapes of a (possibly pruned) ViT/DeiT.

    heads[l] / inter[l] follow the reference's pruned-model conventions:
    head size stays fixed (``modeling/layers/transformer_encoder.py:30``,
    HF ``prune_heads``), FFN width is an arbitrary positive integer
    (``inference_model_patcher.py:152-159``).
    """

    hidden: int = 192
    layers: int = 12
    heads: List[int] = field(default_factory=lambda: [3] * 12)
    inter: List[int] = field(default_factory=lambda: [768] * 12)
    head_size: int = 64
    tokens: int = 197          # 197 ViT/DeiT-as-ViT, 198 DeiT with distillation token
    eps: float = 1e-12         # HF default; 1e-5 in the TF dialect
    gelu: str = "erf"          # "erf" (HF) or "tanh" (modeling/torch_layers/activation.py:4-7)
    num_labels: int = 1000
    image: int = 224
    patch: int = 16

    @staticmethod
    def deit(name: str, **kw) -> "ViTSpec":
        d, h = {"tiny": (192, 3), "small": (384, 6), "base": (768, 12)}[name]
        base = dict(hidden=d, layers=12, heads=[h] * 12, inter=[4 * d] * 12)
        base.update(kw)
        return ViTSpec(**base)

    @property
    def patches(self) -> int:
        return (self.image // self.patch) ** 2

    def matmul_flops(self) -> float:
        """Algorithmic matmul FLOPs per image, SURVEY.md section 8d formula."""
        S, D = self.tokens, self.hidden
        f = 2.0 * self.patches * (3 * self.patch * self.patch) * D
        for h, i in zip(self.heads, self.inter):
            a = h * self.head_size
            f += 3 * 2 * S * D * a + 2 * 2 * S * S * a + 2 * S * a * D + 2 * 2 * S * D * i
        f += 2.0 * D * self.num_labels
        return f
