"""World-size-2 gloo tests (CPU) of the only multi-process logic next to the path: contiguous batch sharding and the
post-loop counter reduce (deit_pruning/src/utils.py:167-168,221-226).  The forward itself has no collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from edgevisiontransformer_b200.eval_loop import reduce_counters, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 4096 + 3
    lo, hi = shard_range(n, rank, world)
    # every rank "evaluates" its shard: correct = number of even indices, loss = mean index
    idx = torch.arange(lo, hi)
    res = {"eval_accuracy": float((idx % 2 == 0).float().mean()), "eval_loss": float(idx.float().mean()), "inference_time": 1.0 + rank}
    out = reduce_counters(dict(res), "cpu")
    q.put((rank, lo, hi, res, out))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_reduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, res0, out0), (r1, lo1, hi1, res1, out1) = got
    assert (lo0, hi0, lo1, hi1) == (0, 2050, 2050, 4099)             # contiguous, disjoint, covering
    for k in res0:                                                    # rank 0 holds sum / world_size
        assert abs(out0[k] - (res0[k] + res1[k]) / 2) < 1e-9


def test_shard_range_edges():
    assert shard_range(10, 0, 4) == (0, 3) and shard_range(10, 3, 4) == (9, 10)
    assert shard_range(2, 3, 4) == (2, 2)                              # more ranks than items: empty shard
    cover = [shard_range(4096, r, 8) for r in range(8)]
    assert cover[0] == (0, 512) and cover[-1] == (3584, 4096)


def test_chunk_schedule_properties():
    """Host logic of the pipelined eval loop: the chunk sizes cover the batch, never exceed the chunk, ramp up by at most 3x
    (every copy hides behind the previous forward) and do not end in a short tail when it can be folded into the ramp."""
    from edgevisiontransformer_b200.eval_loop import chunk_schedule
    for batch in (1, 6, 63, 64, 100, 511, 512, 1000, 1024, 2048, 4095, 4096, 5000):
        for chunk in (4, 64, 256, 512, 1024):
            s = chunk_schedule(batch, chunk)
            assert sum(s) == batch and all(0 < n <= chunk for n in s), (batch, chunk, s)
            assert all(s[i] <= 3 * s[i - 1] for i in range(1, len(s))), (batch, chunk, s)
    assert chunk_schedule(4096, 1024) == [128, 384, 512, 1024, 1024, 1024]
    assert chunk_schedule(512, 512) == [64, 192, 256]


def test_uniform_schedule_properties():
    """Steady-state schedule of `PipelinedClassifier.submit` (the previous batch still runs, nothing to ramp for): full chunks,
    never more than the chunk, and no tail shorter than half a chunk -- a short tail is split evenly with its neighbour."""
    from edgevisiontransformer_b200.eval_loop import uniform_schedule
    for batch in (1, 6, 63, 64, 100, 511, 512, 513, 1000, 1024, 1030, 2048, 2100, 4095, 4096, 5000):
        for chunk in (4, 64, 256, 512, 1024):
            s = uniform_schedule(batch, chunk)
            assert sum(s) == batch and all(0 < n <= chunk for n in s), (batch, chunk, s)
            if len(s) > 1:
                assert min(s) >= chunk // 4, (batch, chunk, s)      # the evenly split tail is at least a quarter chunk
    assert uniform_schedule(4096, 1024) == [1024] * 4
    assert uniform_schedule(2100, 1024) == [1024, 538, 538]
    assert uniform_schedule(1030, 1024) == [515, 515]
    assert uniform_schedule(7, 1024) == [7]
