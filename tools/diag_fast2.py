"""Which part of the fast path moves eye bytes relative to the exact path?  VD3D_FAST_DEBUG bits, one context each."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import dibr as O  # noqa: E402
from visiondepth3d_b200 import _lib  # noqa: E402
from visiondepth3d_b200 import render_3d as R  # noqa: E402
from visiondepth3d_b200.synth import synth_frame  # noqa: E402


def eyes(ctx, w, h, kind, exact):
    _lib._default_ctx[0] = ctx
    ctx.set_exact(exact)
    R.reset_temporal_state()
    fr, dp = synth_frame(1, w, h, kind)
    l, r, s = R.pixel_shift_cuda(O.bgr_to_rgb01(fr), O.depth_bgr_to_01(dp), w, h, 4.5, -1.5, -6.0, blur_ksize=9,
                                 feather_strength=10.0, zero_parallax_strength=0.01)
    return l, r, s.numpy()


for (w, h, kind) in ((1920, 1080, "natural"), (320, 180, "natural")):
    os.environ["VD3D_FAST_DEBUG"] = "0"
    base = _lib.Context(0)
    le, re_, se = eyes(base, w, h, kind, True)
    for bits in (0, 1, 2, 4, 3, 7):
        os.environ["VD3D_FAST_DEBUG"] = str(bits)
        c = _lib.Context(0)
        lf, rf, sf = eyes(c, w, h, kind, False)
        d = np.abs(lf.astype(int) - le.astype(int))
        print(f"{w}x{h} {kind} dbg={bits}: flips {(d > 0).mean():.6f} max {d.max()} shift max|d| {np.abs(sf - se).max():.2e}")
        # where are the flips?  fraction inside the image interior / by feather weight is not observable here; rows:
        if bits == 7:
            ys, xs, _ = np.nonzero(d)
            if len(ys):
                print("    dbg=7 flips rows", np.percentile(ys, [0, 25, 50, 75, 100]), "cols", np.percentile(xs, [0, 25, 50, 75, 100]))
        c.close()
    base.close()
