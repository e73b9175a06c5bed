"""Eval loop around the forward: the immediate caller of the hot path.

Mirrors ``evaluate`` in the reference (deit_pruning/src/utils.py:151-228, are_16_heads/classifier_eval.py:22-108):
DataLoader -> host->device copy -> ``model(images).logits`` -> argmax -> accuracy, with the per-rank counters
summed to rank 0 when distributed.  B200-first changes: pinned host staging with the H2D copy of chunk i+1
overlapped with the forward of chunk i on a second stream, and argmax on the GPU so only B int64 leave the
device per batch instead of B x 1000 floats (the reference does ``logits.cpu().numpy()`` per batch).
"""
from __future__ import annotations

import time
from typing import Dict, Iterable, Optional

import torch


class PipelinedClassifier:
    """Runs a B200ViTForImageClassification over a HOST batch in chunks, double-buffering the H2D copies."""

    def __init__(self, model, chunk: int = 512):
        self.model = model
        self.chunk = int(chunk)
        self.device = next(model.parameters()).device
        # device staging buffers, per pixel storage type: the host batch is copied as it is (f32, bf16 or raw uint8) and the
        # conversion happens inside the library's patch gather, so bf16 / u8 loaders move 2x / 4x fewer bytes over PCIe
        self._bufs_by_dtype = {}
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._ready = [torch.cuda.Event() for _ in range(2)]
        self._free = [torch.cuda.Event() for _ in range(2)]
        self._host_out = None
        self._pinned_out = {}

    @torch.no_grad()
    def _run(self, host_pixels: torch.Tensor, want_logits: bool):
        if host_pixels.is_cuda:
            raise ValueError("PipelinedClassifier takes host tensors; call the model directly for device tensors")
        B = host_pixels.shape[0]
        c = self.model.config
        if host_pixels.dtype not in (torch.float32, torch.bfloat16, torch.uint8):
            host_pixels = host_pixels.float()
        bufs = self._bufs_by_dtype.get(host_pixels.dtype)
        if bufs is None:
            bufs = self._bufs_by_dtype[host_pixels.dtype] = [
                torch.empty((self.chunk, 3, c.image_size, c.image_size), dtype=host_pixels.dtype, device=self.device)
                for _ in range(2)]
        main = torch.cuda.current_stream(self.device)
        out = torch.empty((B, c.num_labels) if want_logits else (B,), dtype=torch.float32 if want_logits else torch.int64,
                          device=self.device)
        # Chunk schedule: the first H2D copy cannot overlap anything, so the chunks ramp up geometrically (chunk/8, /4, /2,
        # then full chunks): every later copy hides behind the forward of the chunk before it (a forward costs ~3x its
        # copy per image), and only the first small copy is exposed.
        bounds, s0 = [], 0
        for n in chunk_schedule(B, self.chunk):
            bounds.append((s0, s0 + n))
            s0 += n
        host_out = None
        if want_logits and host_pixels.is_pinned():
            # pinned result buffer, allocated once per batch size (cudaHostAlloc costs milliseconds); the caller owns the
            # returned tensor until the next call with the same batch size
            host_out = self._pinned_out.get(B)
            if host_out is None:
                host_out = self._pinned_out[B] = torch.empty((B, c.num_labels), dtype=torch.float32).pin_memory()
        for k in range(2):
            self._free[k].record(main)
        for i, (s, e) in enumerate(bounds):
            k = i & 1
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(self._free[k])          # buffer k no longer read by forward i-2
                bufs[k][: e - s].copy_(host_pixels[s:e], non_blocking=True)
                self._ready[k].record(self._copy_stream)
            main.wait_event(self._ready[k])
            lg = self.model(bufs[k][: e - s]).logits
            if want_logits:
                out[s:e] = lg
                if host_out is not None:
                    host_out[s:e].copy_(out[s:e], non_blocking=True)   # D2H of this chunk overlaps the next forward
            else:
                out[s:e] = lg.argmax(dim=-1)
            self._free[k].record(main)
        if host_out is not None:
            self._host_out = host_out
        return out

    def logits(self, host_pixels: torch.Tensor) -> torch.Tensor:
        """[B, num_labels] f32 on the HOST (synchronises)."""
        self._host_out = None
        dev_out = self._run(host_pixels, True)
        if self._host_out is not None:                       # pinned input: chunk-wise async D2H already queued
            torch.cuda.current_stream(self.device).synchronize()
            out, self._host_out = self._host_out, None
            return out.clone() if out.numel() < (1 << 22) else out   # small results: hand out a private copy
        return dev_out.cpu()

    def predict(self, host_pixels: torch.Tensor) -> torch.Tensor:
        """argmax class ids [B] on the HOST; only 8 bytes per image cross PCIe on the way back."""
        return self._run(host_pixels, False).cpu()


def chunk_schedule(batch: int, chunk: int) -> list:
    """Chunk sizes for one host batch: a geometric ramp (chunk/8, /4, /2) up to full chunks, so that only the first, small
    H2D copy is exposed; a short tail (less than half a chunk) is folded into the earliest ramp chunk that can take it
    without outgrowing 3x its predecessor (its copy must still hide behind the previous forward) -- a 128-image forward at
    the end of a 4096-image batch runs well below the large-batch rate."""
    sizes, rem, size = [], batch, max(1, min(chunk, max(64, chunk // 8)))
    while rem > 0:
        n = min(size, rem)
        sizes.append(n)
        rem -= n
        size = min(chunk, size * 2)
    if len(sizes) > 2 and sizes[-1] < chunk // 2:
        tail = sizes[-1]
        for i in range(1, len(sizes) - 1):
            if sizes[i] + tail <= min(chunk, 3 * sizes[i - 1]):
                sizes[i] += tail
                sizes.pop()
                break
    return sizes


def evaluate(eval_data: Iterable, model, eval_batch_size: int = 100, device=None, result: Optional[Dict] = None,
             distributed: bool = False, num_workers: int = 0, chunk: int = 512) -> Dict[str, float]:
    """Drop-in for deit_pruning/src/utils.py:151 ``evaluate`` (same result keys).

    ``eval_data`` is a map-style dataset of (image, label) like the reference's ImageFolder, or any iterable of
    already-batched (images, labels).  ``eval_loss`` keeps the reference's (odd) definition: the mean logit."""
    from torch.utils.data import DataLoader, Dataset, DistributedSampler, SequentialSampler
    import torch.distributed as dist

    if isinstance(eval_data, Dataset):
        sampler = DistributedSampler(eval_data) if distributed else SequentialSampler(eval_data)
        loader = DataLoader(eval_data, sampler=sampler, batch_size=eval_batch_size, num_workers=num_workers, pin_memory=True)
    else:
        loader = eval_data
    runner = PipelinedClassifier(model, chunk=min(chunk, max(1, eval_batch_size)))
    dev = runner.device
    correct = torch.zeros((), dtype=torch.int64, device=dev)
    loss_sum = torch.zeros((), dtype=torch.float64, device=dev)
    n_examples, n_steps, inference_time = 0, 0, 0.0
    for images, labels in loader:
        t0 = time.time()
        lg = runner._run(images if not images.is_cuda else images.cpu(), True)
        pred = lg.argmax(dim=-1)
        correct += (pred == labels.to(dev, non_blocking=True)).sum()
        loss_sum += lg.double().mean()
        torch.cuda.current_stream(dev).synchronize()
        inference_time += time.time() - t0
        n_examples += images.shape[0]
        n_steps += 1
    result = dict(result or {})
    result["eval_loss"] = float(loss_sum.item()) / max(n_steps, 1)
    result["eval_accuracy"] = float(correct.item()) / max(n_examples, 1)
    result["inference_time"] = inference_time
    if distributed and dist.is_available() and dist.is_initialized():
        reduce_counters(result, dev)
    return result


def reduce_counters(result: Dict[str, float], device) -> Dict[str, float]:
    """dist.reduce(SUM)->rank 0 then / world_size for each scalar, as deit_pruning/src/utils.py:221-226.
    The only collective anywhere near the path, and it runs after the loop."""
    import torch.distributed as dist
    keys = sorted(result)
    t = torch.tensor([float(result[k]) for k in keys], dtype=torch.float64,
                     device=device if dist.get_backend() == "nccl" else "cpu")
    dist.reduce(t, 0, op=dist.ReduceOp.SUM)
    t /= dist.get_world_size()
    for k, v in zip(keys, t.tolist()):
        result[k] = v
    return result


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous batch-shard [lo, hi) of rank `rank`."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)
