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
    flat = torch.cat([sd[k].reshape(-1).float() for k in keys])
    if device is not None:
        flat = flat.to(device)
    dist.broadcast(flat, src=src)
    flat = flat.cpu()
    off = 0
    for k in keys:
        n = sd[k].numel()
        sd[k] = flat[off:off + n].reshape(sd[k].shape).to(sd[k].dtype)
        off += n
    return sd
