"""Frame-parallel sharding helpers: contiguous chunks per rank, one
weight broadcast at init, no per-frame collective.  Backend-agnostic torch.distributed
(NCCL on the GPU box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def chunk_range(n_frames: int, rank: int, world: int):
    """Contiguous [start, stop) of frames for `rank`; sizes differ by at most one."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_state_dict(sd, src=0, device=None):
    """One flat broadcast of all floating tensors of `sd` from `src` (in place)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sd
    keys = [k for k, v in sd.items() if torch.is_floating_point(v)]
    if dist.get_rank() == src:
        flat = torch.cat([sd[k].reshape(-1).float() for k in keys])
        if device is not None:
            flat = flat.to(device)
    else:  # receivers only need the shapes (their tensors may live on the meta device)
        flat = torch.empty(sum(sd[k].numel() for k in keys), dtype=torch.float32, device=device or "cpu")
    dist.broadcast(flat, src=src)
    flat = flat.cpu()
    off = 0
    for k in keys:
        n = sd[k].numel()
        sd[k] = flat[off:off + n].reshape(sd[k].shape).to(sd[k].dtype)
        off += n
    return sd


def render_chunk_exact(ctx, frames, depths, rp, render_frame, advance_state, rank=None, world=None,
                        reset_global=False):
    """Exact frame-sharded rendering of one clip.

    Every rank holds the whole list of (frame, depth) pairs or at least its own chunk.  Rank r
    (1) receives the temporal state after frame start-1 from rank r-1, (2) advances it over its own
    chunk without rendering and forwards it to rank r+1 -- this short scalar/plane chain is the only
    sequential part -- then (3) re-imports its start state and renders its chunk.  The result is
    bit-identical to one GPU rendering the clip sequentially.  Like render_sbs_3d, rank 0 starts from fresh per-render
    state but KEEPS the module-singleton trackers (DepthPercentileEMA, ConvergenceEMA, FloatingWindowTracker,
    FloatingBarEaser persist across renders in the reference, core/render_3d.py:284-285,500,511) unless reset_global.
    The chain is sequential: rank r starts rendering after ranks 0..r-1 have advanced over their chunks
    (N_frames x t(k_stats) in total); see DESIGN.md section 7 for what that costs.
    Returns (start, stop, [rendered frames of the chunk])."""
    import numpy as np
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    start, stop = chunk_range(len(frames), rank, world)
    if rank > 0:
        n = torch.zeros(1, dtype=torch.int64)
        dist.recv(n, src=rank - 1)
        blob = torch.empty(int(n.item()), dtype=torch.uint8)
        dist.recv(blob, src=rank - 1)
        start_state = blob.numpy().copy()
        ctx.import_state(start_state)
    else:
        try:
            ctx.reset(3 if reset_global else 2)   # STATE_GLOBAL | STATE_CLIP, or the per-render state only
        except TypeError:                         # contexts without the split (host-logic tests)
            ctx.reset()
        start_state = None
    if rank + 1 < world:
        if start_state is None:
            rank0_state = ctx.export_state()      # what rank 0 returns to after its advance pass (singletons included)
        for i in range(start, stop):
            advance_state(frames[i], depths[i], rp, ctx=ctx)
        end_state = torch.from_numpy(np.ascontiguousarray(ctx.export_state()))
        dist.send(torch.tensor([end_state.numel()], dtype=torch.int64), dst=rank + 1)
        dist.send(end_state, dst=rank + 1)
        if start_state is not None:
            ctx.import_state(start_state)
        else:
            ctx.import_state(rank0_state)
    outs = [render_frame(frames[i], depths[i], rp, ctx=ctx) for i in range(start, stop)]
    return start, stop, outs
