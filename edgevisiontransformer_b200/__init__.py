"""edgevisiontransformer_b200: B200-native (sm_100a) ViT / DeiT / T2T inference forward.

The product is libevt.so (hand-written CUDA behind the C ABI in include/evt.h); this package is the
Python mirror of the reference's module surface for that path.  No CPU fallback anywhere.
"""
from .modeling_vit import (B200ViTConfig, B200ViTForImageClassification, ImageClassifierOutput,  # noqa: F401
                           config_from_state_dict)

__all__ = ["B200ViTConfig", "B200ViTForImageClassification", "ImageClassifierOutput", "config_from_state_dict"]
