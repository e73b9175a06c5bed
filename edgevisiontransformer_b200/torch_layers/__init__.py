"""Drop-in for the reference's ``modeling/torch_layers`` package (same class names, constructor arguments,
parameter names and error behaviour), running on libevt's sm_100a kernels.

    from edgevisiontransformer_b200.torch_layers import Attention, FeedForward, LayerNorm, Residual, gelu

``get_attention`` / ``get_ffn`` mirror ``utils.get_attention / get_ffn(is_tf=False)`` (utils.py:322-365).
"""
from .activation import gelu  # noqa: F401
from .attention import Attention  # noqa: F401
from .ffn import FeedForward  # noqa: F401
from .norm import LayerNorm  # noqa: F401
from .residual import Residual  # noqa: F401


def get_attention(h=768, a=12, h_k=None, n=128):
    """utils.py:322-339 with is_tf=False."""
    return LayerNorm([n, h], Residual(Attention(h, a, h_k)))


def get_ffn(h=768, i=3072, n=128, only_ffn=False):
    """utils.py:342-365 with is_tf=False."""
    if only_ffn:
        return FeedForward(h, i)
    return LayerNorm([n, h], Residual(FeedForward(h, i)))
