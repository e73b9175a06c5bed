"""Eval loop around the forward: the immediate caller of the hot path.

Mirrors ``evaluate`` in the reference (deit_pruning/src/utils.py:151-228, are_16_heads/classifier_eval.py:22-108):
DataLoader -> host->device copy -> ``model(images).logits`` -> argmax -> accuracy, with the per-rank counters
summed to rank 0 when distributed.  B200-first changes: pinned host staging with the H2D copy of chunk i+1
overlapped with the forward of chunk i on a second stream, and argmax on the GPU so only B int64 leave the
device per batch instead of B x 1000 floats (the reference does ``logits.cpu().numpy()`` per batch).
"""
from __future__ import annotations

import contextlib
import os
import time
from typing import Dict, Iterable, Optional

import torch


@contextlib.contextmanager
def near_gpu(device=None):
    """Run the body with this thread bound to the CPUs closest to `device` (NVML's ideal CPU affinity), then restore the
    previous affinity.  Pinned host buffers allocated inside land on the GPU's NUMA node, so their H2D copies do not cross the
    socket interconnect -- on a two-socket box with eight GPUs copying at once that link, not PCIe, is what saturates.
    Yields a description of the binding, or None when NVML / the affinity call is not available (nothing is changed then)."""
    old, note = None, None
    try:
        import pynvml
        idx = torch.device(device if device is not None else torch.cuda.current_device())
        idx = idx.index if idx.index is not None else torch.cuda.current_device()
        prop = torch.cuda.get_device_properties(idx)
        pynvml.nvmlInit()
        try:
            bus = "%08x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        old = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        now = os.sched_getaffinity(0)
        note = "cpus %d-%d (%d of %d)" % (min(now), max(now), len(now), len(old))
    except Exception:
        if old is not None:
            try:
                os.sched_setaffinity(0, old)
            except Exception:
                pass
        old = None
    try:
        yield note
    finally:
        if old is not None:
            try:
                os.sched_setaffinity(0, old)
            except Exception:
                pass


class PendingLogits:
    """Handle of one `PipelinedClassifier.submit` call: the copies and launches are queued, `result()` waits for them."""

    def __init__(self, dev_out, host_out, done):
        self._dev_out, self._host_out, self._done = dev_out, host_out, done

    def done(self) -> bool:
        return self._done.query()

    def result(self) -> torch.Tensor:
        """[B, num_labels] f32 on the HOST.  With pinned input this is one of the runner's two pinned result buffers for the
        batch size: valid until the second-next submit of that size (clone it to keep it longer)."""
        self._done.synchronize()
        return self._host_out if self._host_out is not None else self._dev_out.cpu()


class PipelinedClassifier:
    """Runs a B200ViTForImageClassification over a HOST batch in chunks, double-buffering the H2D copies.

    `logits(host)` is the synchronous call; `submit(host)` queues the same work and returns a handle, so an evaluation loop can
    queue batch i+1 before it collects batch i -- the H2D copies of the next batch then run under the last forwards of the
    current one (the staging buffers are guarded by events, not by a synchronise), which is what a DataLoader-fed loop with
    non-blocking copies does in the reference (deit_pruning/src/utils.py:188-199)."""

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
        self._pinned_sel = {}
        self._last = None              # handle of the most recent submit (steady state = it is still running)
        self._next_buf = 0

    @torch.no_grad()
    def _run(self, host_pixels: torch.Tensor, want_logits: bool, ramp: bool = True):
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
        for n in (chunk_schedule(B, self.chunk) if ramp else uniform_schedule(B, self.chunk)):
            bounds.append((s0, s0 + n))
            s0 += n
        host_out = None
        if want_logits and host_pixels.is_pinned():
            # pinned result buffer, allocated once per batch size (cudaHostAlloc costs milliseconds); the caller owns the
            # returned tensor until the next call with the same batch size
            # two of them, alternating: the result of the previous submit may not have been collected yet
            pair = self._pinned_out.get(B)
            if pair is None:
                with near_gpu(self.device):
                    pair = self._pinned_out[B] = [torch.empty((B, c.num_labels), dtype=torch.float32).pin_memory() for _ in range(2)]
            sel = self._pinned_sel.get(B, 0)
            self._pinned_sel[B] = sel ^ 1
            host_out = pair[sel]
        # (no re-recording of the `free` events here: each was last recorded after the forward that read its buffer, possibly
        #  in the previous call -- waiting for exactly that forward is what lets this call's first copies start early)
        for s, e in bounds:
            # the two staging buffers alternate ACROSS calls too: with one chunk per call (batch <= chunk) a per-call index would
            # put every batch in buffer 0 and the next batch's copy would wait for this batch's forward
            k = self._next_buf
            self._next_buf ^= 1
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

    def submit(self, host_pixels: torch.Tensor) -> PendingLogits:
        """Queue H2D copies, forwards and the D2H of the logits for one host batch; returns without waiting.  At most two
        submits may be outstanding (two pinned result buffers per batch size).  While the previous submit is still running
        the chunks are uniform -- the geometric ramp only exists to shorten the one copy nothing can hide."""
        steady = self._last is not None and not self._last.done()
        self._host_out = None
        dev_out = self._run(host_pixels, True, ramp=not steady)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(self.device))
        h = PendingLogits(dev_out, self._host_out, done)
        self._host_out = None
        self._last = h
        return h

    def logits(self, host_pixels: torch.Tensor) -> torch.Tensor:
        """[B, num_labels] f32 on the HOST (synchronises)."""
        out = self.submit(host_pixels).result()
        return out.clone() if out.numel() < (1 << 22) else out      # small results: hand out a private copy

    def predict(self, host_pixels: torch.Tensor) -> torch.Tensor:
        """argmax class ids [B] on the HOST; only 8 bytes per image cross PCIe on the way back."""
        return self._run(host_pixels, False).cpu()


def uniform_schedule(batch: int, chunk: int) -> list:
    """Full chunks, a short tail (less than half a chunk) folded into the last one's neighbour by splitting evenly."""
    n = max(1, -(-batch // chunk))
    if n > 1 and batch - (n - 1) * chunk < chunk // 2:
        n -= 1                                   # the tail rides on the last full chunk, split in two equal parts below
        base = (n - 1) * chunk
        rest = batch - base
        return [chunk] * (n - 1) + ([rest] if rest <= chunk else [rest - rest // 2, rest // 2])
    return [min(chunk, batch - i * chunk) for i in range(n)]


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
        # Queued, not waited for: the staging buffers are guarded by events, so the copies of the next batch run under the last
        # forwards of this one (and under the loader's work for the batch after).  `inference_time` is the time this loop spends
        # queueing GPU work plus the final drain -- the reference's per-batch wall time (utils.py:190-199) without its per-batch
        # synchronise.
        lg = runner._run(images if not images.is_cuda else images.cpu(), True, ramp=(n_steps == 0))
        pred = lg.argmax(dim=-1)
        correct += (pred == labels.to(dev, non_blocking=True)).sum()
        loss_sum += lg.double().mean()
        inference_time += time.time() - t0
        n_examples += images.shape[0]
        n_steps += 1
    t0 = time.time()
    torch.cuda.current_stream(dev).synchronize()
    inference_time += time.time() - t0
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
