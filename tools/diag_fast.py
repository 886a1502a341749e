"""fast vs exact DIBR arithmetic on the same inputs: shift map, scalars, eyes (diagnostic for the parity budget)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import dibr as O  # noqa: E402
from visiondepth3d_b200 import render_3d as R  # noqa: E402
from visiondepth3d_b200.synth import synth_frame  # noqa: E402

ctx = R._ctx()
for (w, h, kind) in ((320, 180, "natural"), (1920, 1080, "natural"), (3840, 2160, "natural"), (320, 180, "smooth")):
    res = {}
    for mode in ("exact", "fast"):
        ctx.set_exact(mode == "exact")
        R.reset_temporal_state()
        out = []
        for i in range(2):
            fr, dp = synth_frame(i, w, h, kind)
            infos = []
            l, r, s = R.pixel_shift_cuda(O.bgr_to_rgb01(fr), O.depth_bgr_to_01(dp), w, h, 4.5, -1.5, -6.0, blur_ksize=9,
                                         feather_strength=10.0, zero_parallax_strength=0.01, _info=infos)
            out.append((l, r, s.numpy(), infos[0]))
        res[mode] = out
    for i in range(2):
        le, re_, se, ie = res["exact"][i]
        lf, rf, sf, if_ = res["fast"][i]
        d = np.abs(lf.astype(int) - le.astype(int))
        print(f"{w}x{h} {kind} frame {i}: shift max|d| {np.abs(sf - se).max():.3e} mean|d| {np.abs(sf - se).mean():.3e} "
              f"(|shift| max {np.abs(se).max():.4f}); eye max {d.max()} flips {(d > 0).mean():.5f}")
        for f in ("subj_raw", "stretch_lo", "stretch_hi", "subj_shaped", "zero_parallax_offset"):
            a, b = getattr(ie, f), getattr(if_, f)
            print(f"    {f}: exact {a:.9g} fast {b:.9g} diff {abs(a - b):.3e}")
ctx.set_exact(False)
