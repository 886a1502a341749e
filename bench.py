#!/usr/bin/env python
"""bench.py -- end-to-end frames/s of the depth -> stereo hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 1080p|4k] [--impl ours|reference]

A step = one batch of `frames_per_step` synthetic frames through the hot path
(depth forward when a depth engine is built + the DIBR frame loop of render_sbs_3d).
`value`  : frames/s, inputs resident in HBM, timed with CUDA events on the engine's stream.
`e2e`    : the same through the public host-buffer API (vd3d_render_clip, pinned host
           frames in, packed frames out, H2D/D2H inside the timed region).
`roofline`: dominant DIBR kernel (compose) and the whole DIBR stage vs measured HBM peak.
`cpu_baseline`: the oracle port timed on a bounded sample on the host cores (rank 0, N=1).
`--impl reference`: the CPU port of the reference path, same JSON line.
Multi-GPU: one process per GPU (torchrun), contiguous chunks of frames per rank, no
per-frame collective; max-over-ranks device time.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]
    "1080p": dict(name="1080p synthetic clip, Depth-Anything-V2-Base, Half-SBS", w=1920, h=1080,
                  fmt="Half-SBS", preserve=False, model="vitb", pool=24, frames_per_step=8,
                  # 3*Ws*Hs + 1*Ws*Hs + 8*Wt*Ht + 3*Wout*Hout
                  dibr_bytes=3 * 1920 * 1080 + 1920 * 1080 + 8 * 960 * 540 + 3 * 1920 * 1080),
    # BASELINE.json configs[2] (per-GPU share)
    "4k": dict(name="4K synthetic clip, Depth-Anything-V2-Large, Full-SBS", w=3840, h=2160,
               fmt="Full-SBS", preserve=True, model="vitl", pool=6, frames_per_step=6,
               dibr_bytes=18 * 3840 * 2160),
}
COMMON = dict(fg=4.5, mg=-1.5, bg=-6.0, sharp=0.2, feather=10.0, ksize=9, tracking=True, floating=True,
              zps=0.01, dof=0.0)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1400.0, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_pool(wl, n, seed0=0):
    """n (frame, depth) pairs of the noise (throughput) set, depth as 3-channel BGR like a decoded depth video."""
    from visiondepth3d_b200.synth import synth_frame
    return [synth_frame(seed0 + i, wl["w"], wl["h"], "noise") for i in range(n)]


def render_params(R, wl, depth_channels=3):
    return R.make_render_params(
        wl["w"], wl["h"], COMMON["fg"], COMMON["mg"], COMMON["bg"], COMMON["sharp"], wl["fmt"], 16 / 9,
        COMMON["dof"], COMMON["feather"], COMMON["ksize"], COMMON["tracking"], COMMON["floating"],
        preserve_original_aspect=wl["preserve"], zero_parallax_strength=COMMON["zps"])


def oracle_params(wl):
    from oracle import dibr as O
    return O.RenderParams(output_width=wl["w"], output_height=wl["h"], fg_shift=COMMON["fg"], mg_shift=COMMON["mg"],
                          bg_shift=COMMON["bg"], sharpness_factor=COMMON["sharp"], output_format=wl["fmt"],
                          dof_strength=COMMON["dof"], feather_strength=COMMON["feather"], blur_ksize=COMMON["ksize"],
                          use_subject_tracking=COMMON["tracking"], use_floating_window=COMMON["floating"],
                          preserve_original_aspect=wl["preserve"], zero_parallax_strength=COMMON["zps"])


def cpu_port_fps(wl, seconds_budget=20.0, max_frames=4):
    """Time the oracle port (numpy, one thread) on a bounded sample of the same workload."""
    from oracle import dibr as O
    from visiondepth3d_b200.synth import synth_frame
    gs, cs = O.GlobalState(), O.ClipState()
    rp = oracle_params(wl)
    fr, dp = synth_frame(0, wl["w"], wl["h"], "noise")
    O.render_frame(gs, cs, fr, dp, rp)  # warm-up frame (first-frame state init)
    n, t0 = 0, time.perf_counter()
    while n < max_frames and (time.perf_counter() - t0) < seconds_budget:
        fr, dp = synth_frame(n + 1, wl["w"], wl["h"], "noise")
        O.render_frame(gs, cs, fr, dp, rp)
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, n


def run_reference(args, wl, rank, world):
    """--impl reference: the reference's CPU path (port: oracle/dibr.py; /root/reference does not
    exist on the GPU box and its Python cannot travel).  Rank 0 only."""
    if rank != 0:
        return
    t_all = time.perf_counter()
    per_step = []
    total = 0
    for s in range(args.warmup + args.steps):
        fps, n = cpu_port_fps(wl, seconds_budget=max(5.0, 60.0 / (args.warmup + args.steps)), max_frames=2)
        if s >= args.warmup:
            per_step.append(fps)
            total += n
    value = float(np.mean(per_step))
    line = {
        "impl": "reference", "metric": "end-to-end frames/sec (depth+stereo)", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "stage": "DIBR frame loop (depth forward not in the CPU arm yet)"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": 1, "kind": "port",
                         "sample": f"{total} frames of the workload, numpy oracle port, 1 thread "
                                   f"(host has {os.cpu_count()} cores)"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libvd3d has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from visiondepth3d_b200 import _lib
    from visiondepth3d_b200 import render_3d as R
    ctx = _lib.Context(local_rank)
    lib = ctx.lib
    rp = render_params(R, wl)
    pl = R.plan_sizes(wl["w"], wl["h"], rp)
    oshape = R.output_shape(rp, pl)
    B, P = wl["frames_per_step"], wl["pool"]
    pool = make_pool(wl, P, seed0=rank * 1000)

    # ---- device-resident inputs / outputs (pool larger than L2: 126 MB) ----
    dev = torch.device("cuda", local_rank)
    f_dev = [torch.from_numpy(f).to(dev) for f, _ in pool]
    d_dev = [torch.from_numpy(d).to(dev) for _, d in pool]
    o_dev = [torch.empty(oshape, dtype=torch.uint8, device=dev) for _ in range(P)]
    in_bytes = sum(t.numel() for t in f_dev) + sum(t.numel() for t in d_dev)
    # ---- pinned host buffers for the end-to-end arm ----
    f_host = [torch.from_numpy(f).pin_memory() for f, _ in pool]
    d_host = [torch.from_numpy(d).pin_memory() for _, d in pool]
    o_host = [torch.empty(oshape, dtype=torch.uint8).pin_memory() for _ in range(P)]
    torch.cuda.synchronize()

    def ptr_array(ts, idx):
        return (C.c_void_p * len(idx))(*[ts[i].data_ptr() for i in idx])

    stream = torch.cuda.ExternalStream(lib.vd3d_stream(ctx.h), device=dev)
    step_idx = [0]

    def step(mem):
        i0 = (step_idx[0] * B) % P
        idx = [(i0 + k) % P for k in range(B)]
        step_idx[0] += 1
        if mem == _lib.MEM_DEVICE:
            a, b, c = ptr_array(f_dev, idx), ptr_array(d_dev, idx), ptr_array(o_dev, idx)
        else:
            a, b, c = ptr_array(f_host, idx), ptr_array(d_host, idx), ptr_array(o_host, idx)
        ctx.check(lib.vd3d_render_clip(ctx.h, B, a, b, 3, wl["h"], wl["w"], C.byref(rp), c, mem, None))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ================= device-resident arm (value) =================
    ctx.reset()
    for _ in range(args.warmup):
        step(_lib.MEM_DEVICE)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    lib.vd3d_profile(ctx.h, 1)
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step(_lib.MEM_DEVICE)
    e1.record(stream)
    barrier()
    ms = reduce_max(e0.elapsed_time(e1))
    launches = ctx.launches - l0
    tot0, n0, tot1, n1 = C.c_double(), C.c_int(), C.c_double(), C.c_int()
    lib.vd3d_profile_collect(ctx.h, 0, C.byref(tot0), C.byref(n0))
    lib.vd3d_profile_collect(ctx.h, 1, C.byref(tot1), C.byref(n1))
    lib.vd3d_profile(ctx.h, 0)
    frames = args.steps * B
    value = world * frames / (ms / 1000.0)

    # ================= end-to-end arm (host buffers through the public API) =================
    ctx.reset()
    for _ in range(args.warmup):
        step(_lib.MEM_HOST)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(_lib.MEM_HOST)
    barrier()
    e2e_s = reduce_max(time.perf_counter() - t0)
    clk = clocks.stop()
    e2e_value = world * frames / e2e_s
    h2d = B * (wl["w"] * wl["h"] * 3 * 2)
    d2h = B * int(np.prod(oshape))

    if rank == 0:
        hbm_peak, tf_peak, which = peaks()
        px = wl["w"] * wl["h"] if wl["preserve"] else pl.resized_width * pl.resized_height
        comp_bytes = 3 * px + 6 * px  # compose: RGB in (u8) + two u8 eyes out
        comp_ms = tot1.value / max(n1.value, 1)
        stage_ms = tot0.value / max(n0.value, 1)
        comp_gbs = comp_bytes / (comp_ms * 1e-3) / 1e9 if comp_ms > 0 else 0.0
        stage_gbs = wl["dibr_bytes"] / (stage_ms * 1e-3) / 1e9 if stage_ms > 0 else 0.0
        line = {
            "metric": "end-to-end frames/sec (depth+stereo)", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": wl["name"], "frames_per_step": B, "frame_pool": P,
                "l2_policy": f"inputs larger than L2 ({in_bytes / 1e6:.0f} MB pool cycled)",
                "depth_model": None, "stage": "DIBR frame loop only (depth forward engine not built yet)",
                "params": COMMON, "sharding": "contiguous chunks per rank, independent temporal state per chunk",
            },
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_compose", "achieved": comp_gbs, "peak": hbm_peak,
                         "unit": "GB/s", "frac": comp_gbs / hbm_peak, "traffic": None, "peak_source": which,
                         "algorithmic_bytes_per_launch": comp_bytes, "avg_launch_ms": comp_ms},
            "roofline_stage": {"bound": "hbm", "what": "whole DIBR frame (ingest..pack)", "achieved": stage_gbs,
                               "peak": hbm_peak, "unit": "GB/s", "frac": stage_gbs / hbm_peak,
                               "algorithmic_bytes_per_frame": wl["dibr_bytes"], "avg_frame_ms": stage_ms},
        }
        if world == 1 and not args.no_cpu_baseline:
            fps, n = cpu_port_fps(wl)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": 1, "kind": "port",
                                    "sample": f"{n} frames of the workload after 1 warm-up frame, numpy oracle port "
                                              f"(DIBR loop), 1 thread of {os.cpu_count()} host cores"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
