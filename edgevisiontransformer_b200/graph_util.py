"""CUDA-graph replay of a Python-composed forward (Swin, T2T): the launch sequence for one input shape is captured once
and replayed, which removes the per-kernel launch and ctypes overhead on the small-batch latency path -- the same
treatment ``B200ViTForImageClassification.forward_graphed`` gives the C++ runtime's forward."""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import torch


def graphed_call(cache: Dict[Tuple[int, ...], tuple], x: torch.Tensor, run: Callable[[torch.Tensor], torch.Tensor],
                 device: torch.device) -> torch.Tensor:
    """``run(static_input) -> output`` captured per input shape; returns a private copy of the output."""
    key = tuple(x.shape)
    ent = cache.get(key)
    with torch.cuda.device(device):
        if ent is None:
            static_in = torch.empty(key, dtype=torch.float32, device=device)
            static_in.copy_(x)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):              # eager warm-up: persistent buffers are allocated outside the graph pool
                for _ in range(2):
                    run(static_in)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_out = run(static_in)
            ent = cache[key] = (g, static_in, static_out)
        g, static_in, static_out = ent
        static_in.copy_(x)
        g.replay()
        return static_out.clone()
