import torch
import torch.nn as nn

from .. import ops
from ._packed import PackedWeights


class FeedForward(nn.Module):
    """modeling/torch_layers/ffn.py:7-17: linear2(gelu_tanh(linear1(x))); the GELU runs in the FC1 GEMM epilogue."""

    def __init__(self, hidden_size, intermediate_size):
        super().__init__()
        self.linear1 = nn.Linear(hidden_size, intermediate_size)
        self.linear2 = nn.Linear(intermediate_size, hidden_size)
        self._packed = PackedWeights()

    def _weights(self):
        ps = [self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias]

        def build():
            i = self.linear1.out_features
            ld = (i + 7) // 8 * 8          # TMA wants 16-byte rows: pad the K dim of linear2 with zero columns
            w2 = torch.zeros((self.linear2.out_features, ld), dtype=torch.bfloat16, device=self.linear2.weight.device)
            w2[:, :i] = self.linear2.weight.to(torch.bfloat16)
            return (self.linear1.weight.to(torch.bfloat16).contiguous(), self.linear1.bias.float().contiguous(), w2,
                    self.linear2.bias.float().contiguous(), ld)
        return self._packed.get(ps, build)

    @torch.no_grad()
    def forward(self, x, residual=None):
        if not x.is_cuda:
            raise RuntimeError("edgevisiontransformer_b200.torch_layers.FeedForward needs CUDA tensors (no CPU fallback)")
        shape = x.shape
        h = shape[-1]
        w1, b1, w2, b2, ld = self._weights()
        if h % 8:
            raise NotImplementedError("hidden_size must be a multiple of 8 (16-byte rows for TMA)")
        xb = ops.cast_bf16(x.float().reshape(-1, h))
        i = w1.shape[0]
        mid = torch.empty((xb.shape[0], ld), dtype=torch.bfloat16, device=x.device)
        ops.linear(xb, w1, b1, act="gelu_tanh", out=mid, n=i)
        res = None if residual is None else residual.float().reshape(-1, h).contiguous()
        out = ops.linear(mid, w2, b2, residual=res, out_dtype=torch.float32, k=i)
        return out.view(*shape)
