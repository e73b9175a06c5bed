"""tanh-GELU of modeling/torch_layers/activation.py:4-7.  Inside FeedForward it is fused into the FC1 GEMM
epilogue; this free function exists for API parity and runs the same formula with torch elementwise ops."""
import math

import torch


def gelu(x):
    cdf = 0.5 * (1.0 + torch.tanh(math.sqrt(2 / math.pi) * (x + 0.044715 * torch.pow(x, 3))))
    return x * cdf
