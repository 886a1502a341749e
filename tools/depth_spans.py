"""Warm device time of every launch class of the batched DA-V2 forward (eager, CUDA events around each launch):
   python tools/depth_spans.py [vitb|vitl] [batch] [reps]      -> per-class us per frame, averaged over reps."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

from visiondepth3d_b200.depth_engine import DepthEngine  # noqa: E402
import torch  # noqa: E402
from transformers import DepthAnythingForDepthEstimation  # noqa: E402

from visiondepth3d_b200.depth_weights import hf_config  # noqa: E402
from visiondepth3d_b200.synth import synth_frame  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "vitb"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
e = DepthEngine(arch, 518, 924)
torch.manual_seed(0)
e.load_state_dict(DepthAnythingForDepthEstimation(hf_config(arch)).eval().state_dict())
frames = [synth_frame(i, 1920, 1080, "natural")[0] for i in range(B)]
for _ in range(3):
    e.infer_batch(frames)
e.check(e.lib.vd3d_depth_profile(e.h, 2))
for _ in range(reps):
    e.infer_batch(frames)
buf = C.create_string_buffer(1 << 16)
e.check(e.lib.vd3d_depth_profile_spans(e.h, buf, len(buf)))
e.check(e.lib.vd3d_depth_profile(e.h, 0))
tot = 0.0
print(f"{arch} batch {B}, {reps} reps: us per frame by launch class")
for line in buf.value.decode().splitlines():
    tag, ms, n = line.rsplit(" ", 2)
    us = float(ms) * 1e3 / (reps * B)
    tot += us
    print(f"  {tag:28s} {us:9.1f} us/frame   ({int(n) // reps} spans per forward, {float(ms) * 1e3 / int(n):8.1f} us each)")
print(f"  {'sum':28s} {tot:9.1f} us/frame")
