"""Exact frame sharding on real GPUs (bench.py --sharding exact, tests/test_sharding_gpu.py).

One clip = the concatenation of the ranks' chunks.  Every rank r
  1. runs the depth forward over its own chunk (no dependency on any other rank; u8 depth stays in HBM),
  2. receives the temporal state after frame start-1 from rank r-1 (one NCCL point-to-point message: the DevState
     struct + the two f32 planes, 4 MB at 1080p), imports it,
  3. advances the state over its chunk WITHOUT rendering (k_stats only: every tracker of the reference + the temporal
     depth plane) and sends the end state to rank r+1 -- this chain is the only sequential part,
  4. re-imports its start state and renders its chunk (vd3d_render_clip on device-resident frames and depths).
The frames are bit-identical to one GPU rendering the whole clip in order.  There is no per-frame collective; the
weights are broadcast once at init.  What the chain costs is reported as `chain_ms`.
"""
import ctypes as C
import json
import time

import numpy as np


def _ptrs(ts):
    return (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


class ExactShard:
    def __init__(self, ctx, deng, rp, src_h, src_w, rank, world, dist, device):
        import torch
        from . import _lib
        self.torch, self._lib, self.ctx, self.deng, self.rp = torch, _lib, ctx, deng, rp
        self.h, self.w, self.rank, self.world, self.dist, self.dev = src_h, src_w, rank, world, dist, device
        lib = ctx.lib
        vp, i = C.c_void_p, C.c_int
        lib.vd3d_depth_infer_batch_device.argtypes = [vp, i, C.POINTER(vp), i, i, C.POINTER(vp), C.POINTER(vp), i]
        lib.vd3d_depth_infer_batch_device.restype = i

    def depth_chunk(self, frames_dev, batch=4):
        """u8 depth [h, w] per frame, device resident (one batched forward per `batch` frames)."""
        torch, lib = self.torch, self.ctx.lib
        depths = [torch.empty((self.h, self.w), dtype=torch.uint8, device=self.dev) for _ in frames_dev]
        for i0 in range(0, len(frames_dev), batch):
            fs, ds = frames_dev[i0:i0 + batch], depths[i0:i0 + batch]
            self.deng.check(lib.vd3d_depth_infer_batch_device(self.deng.h, len(fs), _ptrs(fs), self.h, self.w, _ptrs(ds),
                                                              None, 0))
        return depths

    def _export(self):
        torch, lib, ctx = self.torch, self.ctx.lib, self.ctx
        n = int(lib.vd3d_state_bytes(ctx.h))
        blob = torch.empty(n, dtype=torch.uint8, device=self.dev)
        ctx.check(lib.vd3d_export_state(ctx.h, blob.data_ptr(), n, self._lib.MEM_DEVICE))
        return blob

    def _import(self, blob):
        self.ctx.check(self.ctx.lib.vd3d_import_state(self.ctx.h, blob.data_ptr(), blob.numel(), self._lib.MEM_DEVICE))

    def render_chunk(self, frames_dev, depths_dev, outs_dev):
        """Steps 2-4 of the module docstring.  Returns the seconds this rank spent in the sequential chain."""
        torch, lib, ctx, dist, _lib = self.torch, self.ctx.lib, self.ctx, self.dist, self._lib
        t_chain = 0.0
        start = None
        if self.rank > 0:
            n = torch.zeros(1, dtype=torch.int64, device=self.dev)
            dist.recv(n, src=self.rank - 1)
            start = torch.empty(int(n.item()), dtype=torch.uint8, device=self.dev)
            dist.recv(start, src=self.rank - 1)
            self._import(start)
        else:
            ctx.reset(self._lib.STATE_CLIP)   # a fresh render; the module-singleton trackers persist like the reference's
        if self.rank + 1 < self.world:
            t0 = time.perf_counter()
            if start is None:
                start0 = self._export()
            for f, d in zip(frames_dev, depths_dev):
                ctx.check(lib.vd3d_advance_state(ctx.h, f.data_ptr(), d.data_ptr(), 1, self.h, self.w, C.byref(self.rp),
                                                 _lib.MEM_DEVICE))
            end = self._export()
            dist.send(torch.tensor([end.numel()], dtype=torch.int64, device=self.dev), dst=self.rank + 1)
            dist.send(end, dst=self.rank + 1)
            t_chain = time.perf_counter() - t0
            self._import(start if start is not None else start0)
        n = len(frames_dev)
        ctx.check(lib.vd3d_render_clip(ctx.h, n, _ptrs(frames_dev), _ptrs(depths_dev), 1, self.h, self.w,
                                       C.byref(self.rp), _ptrs(outs_dev), _lib.MEM_DEVICE, None))
        return t_chain


def run(args, wl, rank, local_rank, world, dist, common_config):
    """bench.py --sharding exact: weak scaling, one chunk of steps x frames_per_step frames per rank."""
    import torch
    import bench
    from . import _lib
    from . import render_3d as R
    from .depth_engine import DepthEngine, processed_size
    from .depth_weights import hf_config
    from .sharding import broadcast_state_dict
    from transformers import DepthAnythingForDepthEstimation
    dev = torch.device("cuda", local_rank)
    ctx = _lib.Context(local_rank)
    if rank == 0:
        torch.manual_seed(0)
        with torch.device("cpu"):
            sd = DepthAnythingForDepthEstimation(hf_config(wl["model"])).eval().state_dict()
    else:
        with torch.device("meta"):
            sd = DepthAnythingForDepthEstimation(hf_config(wl["model"])).state_dict()
    sd = broadcast_state_dict(sd, src=0, device=dev)
    ih, iw = processed_size(wl["w"], wl["h"])
    deng = DepthEngine(wl["model"], ih, iw, ctx=ctx)
    deng.load_state_dict(sd)
    del sd
    rp = bench.render_params(R, wl)
    pl = R.plan_sizes(wl["w"], wl["h"], rp)
    oshape = R.output_shape(rp, pl)
    n = args.steps * wl["frames_per_step"]
    P = wl["pool"]
    pool = [torch.from_numpy(f).to(dev) for f, _ in bench.make_pool(wl, P, seed0=rank * 1000)]
    outs_pool = [torch.empty(oshape, dtype=torch.uint8, device=dev) for _ in range(P)]
    frames = [pool[i % P] for i in range(n)]
    outs = [outs_pool[i % P] for i in range(n)]
    sh = ExactShard(ctx, deng, rp, wl["h"], wl["w"], rank, world, dist, dev)

    def one_pass():
        depths = sh.depth_chunk(frames)
        return sh.render_chunk(frames, depths, outs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    one_pass()          # warm-up: workspaces, graphs, NCCL channels
    barrier()
    t0 = time.perf_counter()
    chain = one_pass()
    barrier()
    t = torch.tensor([time.perf_counter() - t0, chain], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        sec = float(t[0])
        line = {
            "metric": "end-to-end frames/sec (depth+stereo)", "value": world * n / sec, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * sec / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16 GEMM operands, fp32 accumulate; DIBR fp32",
            "data": "synthetic", "config": common_config(wl, world, "exact"),
            "run": {"mode": "exact sharding: depth per chunk, state chain rank to rank (k_stats only), render per chunk",
                    "clip_frames": world * n, "chain_ms_max_rank": 1000.0 * float(t[1]),
                    "timed_region_s": sec, "timing": "host clock around barrier + cudaDeviceSynchronize, max over ranks"},
        }
        print(json.dumps(line), flush=True)
