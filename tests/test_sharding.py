"""CPU, gloo, world_size 2: the N>1 host logic (chunking + the single weight broadcast)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from visiondepth3d_b200.sharding import broadcast_state_dict, chunk_range


def test_chunks_partition_the_clip():
    for n in (0, 1, 7, 300, 1000):
        for world in (1, 2, 3, 8):
            spans = [chunk_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)  # different weights per rank before the broadcast
    sd = {"a.weight": torch.randn(5, 3), "b.bias": torch.randn(7), "steps": torch.tensor(rank)}
    broadcast_state_dict(sd, src=0)
    torch.manual_seed(0)
    ref = {"a.weight": torch.randn(5, 3), "b.bias": torch.randn(7)}
    ok = torch.equal(sd["a.weight"], ref["a.weight"]) and torch.equal(sd["b.bias"], ref["b.bias"])
    ok = ok and int(sd["steps"]) == rank  # non-float entries untouched
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and float(t) == float(world)
    q.put((rank, ok, chunk_range(11, rank, world)))
    dist.destroy_process_group()


def test_broadcast_and_reduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert [s for _, _, s in res] == [(0, 6), (6, 11)]


class _FakeCtx:
    """Stand-in for a vd3d context: one EMA scalar as 'temporal state' (host logic test only)."""

    def __init__(self):
        self.s = None

    def reset(self):
        self.s = None

    def export_state(self):
        import numpy as np
        return np.frombuffer(np.float64(-1.0 if self.s is None else self.s).tobytes(), dtype=np.uint8).copy()

    def import_state(self, blob):
        import numpy as np
        v = float(np.frombuffer(np.asarray(blob, dtype=np.uint8).tobytes(), dtype=np.float64)[0])
        self.s = None if v < 0 else v


def _fake_step(ctx, x):
    ctx.s = x if ctx.s is None else 0.9 * ctx.s + 0.1 * x
    return ctx.s


def _exact_worker(rank, world, port, q):
    from visiondepth3d_b200.sharding import render_chunk_exact
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = [float(i * i % 7) for i in range(11)]
    ctx = _FakeCtx()
    render = lambda f, d, rp, ctx: _fake_step(ctx, f) + d  # noqa: E731
    advance = lambda f, d, rp, ctx: _fake_step(ctx, f)     # noqa: E731
    start, stop, outs = render_chunk_exact(ctx, frames, [0.5] * 11, None, render, advance)
    q.put((rank, start, stop, outs))
    dist.destroy_process_group()


def test_exact_sharding_chain_gloo_world3():
    """Host logic of the 8(e) state chain: 3 ranks must reproduce the sequential result exactly."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + (os.getpid() % 500)
    procs = [ctx.Process(target=_exact_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    seq_ctx, seq = _FakeCtx(), []
    for i in range(11):
        seq.append(_fake_step(seq_ctx, float(i * i % 7)) + 0.5)
    got = []
    for _, start, stop, outs in res:
        assert len(outs) == stop - start
        got += outs
    assert got == seq
