import torch
import torch.nn as nn

from .. import ops


class LayerNorm(nn.Module):
    """modeling/torch_layers/norm.py:4-15.  ``input_shape`` may be [n, h] (the reference's op-level models,
    utils.py:338,364: statistics over BOTH trailing dims, affine [n, h]) or a single int / [h]."""

    def __init__(self, input_shape, sub_layer, is_pre=False) -> None:
        super().__init__()
        self.layer_norm = nn.LayerNorm(input_shape)
        self.sub_layer = sub_layer
        self.is_pre = is_pre

    def _norm(self, x):
        if not x.is_cuda:
            raise RuntimeError("edgevisiontransformer_b200.torch_layers.LayerNorm needs CUDA tensors (no CPU fallback)")
        ln = self.layer_norm
        nd = len(ln.normalized_shape)
        if tuple(x.shape[-nd:]) != tuple(ln.normalized_shape):
            raise RuntimeError(f"Given normalized_shape={list(ln.normalized_shape)}, expected input with shape "
                               f"[*, {', '.join(map(str, ln.normalized_shape))}], but got input of size{list(x.shape)}")
        x = x.float().contiguous()
        if nd == 1:
            return ops.layernorm(x, ln.weight.float(), ln.bias.float(), ln.eps, out_dtype=torch.float32)
        lead = x.shape[:-nd]
        y = ops.layernorm2d(x.reshape(-1, *ln.normalized_shape), ln.weight.float(), ln.bias.float(), ln.eps)
        return y.view(*lead, *ln.normalized_shape)

    @torch.no_grad()
    def forward(self, x):
        if self.is_pre:
            return self.sub_layer(self._norm(x))
        return self._norm(self.sub_layer(x))
