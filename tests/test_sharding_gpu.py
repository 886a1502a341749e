"""GPU, >= 2 devices: exact frame sharding over NCCL (visiondepth3d_b200/sharding_bench.py) is bit-identical to one
GPU rendering the clip in order.  Skipped on a one-GPU box (the gloo / single-GPU two-context tests cover the logic)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H, N = 320, 180, 14


def _rp(R):
    return R.make_render_params(W, H, 4.5, -1.5, -6.0, 0.2, "Half-SBS", 16 / 9, 0.0, 10.0, 9, True, True,
                                zero_parallax_strength=0.01)


def _worker(rank, world, port, outdir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from visiondepth3d_b200 import _lib
    from visiondepth3d_b200 import render_3d as R
    from visiondepth3d_b200.sharding import chunk_range
    from visiondepth3d_b200.sharding_bench import ExactShard
    from visiondepth3d_b200.synth import synth_frame
    dev = torch.device("cuda", rank)
    ctx = _lib.Context(rank)
    rp = _rp(R)
    pl = R.plan_sizes(W, H, rp)
    a, b = chunk_range(N, rank, world)
    pairs = [synth_frame(i, W, H, "natural") for i in range(a, b)]
    frames = [torch.from_numpy(f).to(dev) for f, _ in pairs]
    depths = [torch.from_numpy(np.ascontiguousarray(d[..., 0])).to(dev) for _, d in pairs]
    outs = [torch.empty(R.output_shape(rp, pl), dtype=torch.uint8, device=dev) for _ in pairs]
    sh = ExactShard(ctx, None, rp, H, W, rank, world, dist, dev)
    sh.render_chunk(frames, depths, outs)
    torch.cuda.synchronize()
    np.save(os.path.join(outdir, f"rank{rank}.npy"), np.stack([o.cpu().numpy() for o in outs]))
    dist.barrier()
    dist.destroy_process_group()


def test_exact_sharding_over_nccl_is_bit_identical(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world = 2
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from visiondepth3d_b200 import render_3d as R
    from visiondepth3d_b200.synth import synth_frame
    rp = _rp(R)
    R.reset_temporal_state()
    seq = [R.render_frame(f, np.ascontiguousarray(d[..., 0]), rp) for f, d in (synth_frame(i, W, H, "natural") for i in range(N))]
    got = np.concatenate([np.load(os.path.join(str(tmp_path), f"rank{r}.npy")) for r in range(world)])
    assert got.shape[0] == N
    for i in range(N):
        assert np.array_equal(got[i], seq[i]), i
