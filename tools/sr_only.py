"""Real-ESRGAN stage only (SRVGGNetCompact x4, random-init realesr-general-x4v3 shape) on device-resident frames.
   python tools/sr_only.py [w h] [num_conv] [frames]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from visiondepth3d_b200 import _lib  # noqa: E402
from visiondepth3d_b200 import merged_pipeline as MP  # noqa: E402
from visiondepth3d_b200.synth import synth_frame  # noqa: E402

w = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
h = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
nc = int(sys.argv[3]) if len(sys.argv) > 3 else 32
n = int(sys.argv[4]) if len(sys.argv) > 4 else 8
eng = MP.SrEngine(MP.random_srvgg_state_dict(nc, 0))
dev = torch.device("cuda", 0)
f = torch.from_numpy(synth_frame(1, w, h, "natural")[0]).to(dev)
o = torch.empty((4 * h, 4 * w, 3), dtype=torch.uint8, device=dev)
gflop = 2.0 * w * h * 9 * 64 * (64 * (nc + 1) + 48) / 1e9     # with the first conv's input padded to 64 channels
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        eng.check(eng.lib.vd3d_sr_forward(eng.h, f.data_ptr(), h, w, eng.num_conv, o.data_ptr(), _lib.MEM_DEVICE))
    eng.ctx.check(eng.lib.vd3d_sync(eng.ctx.h))
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"{w}x{h} -> {4 * w}x{4 * h}, {nc} body convs: {dt * 1e3:.2f} ms/frame = {1 / dt:.1f} frames/s, {gflop / dt / 1e3:.0f} TFLOP/s ({gflop:.0f} GFLOP/frame)")
