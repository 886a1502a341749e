"""Deterministic synthetic clips; numpy only.

smooth set: ramps + moving disc (parity gate);  noise set: white-noise RGB
(throughput / adversarial LSB case).  Frame i uses default_rng(1000+i).
"""
import numpy as np


def synth_frame(i: int, w: int, h: int, kind: str = "smooth"):
    """Return (frame_bgr u8 [h,w,3], depth_bgr u8 [h,w,3]) for frame index i."""
    rng = np.random.default_rng(1000 + i)
    y, x = np.mgrid[0:h, 0:w]
    cx = (w // 4 + 20 * i) % w
    cy = h // 2
    disc = (x - cx) ** 2 + (y - cy) ** 2 <= (h // 6) ** 2
    if kind == "noise":
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    else:
        rgb = np.empty((h, w, 3), dtype=np.uint8)
        rgb[..., 0] = (x * 255 // max(w - 1, 1)).astype(np.uint8)
        rgb[..., 1] = (y * 255 // max(h - 1, 1)).astype(np.uint8)
        rgb[..., 2] = ((x + y + 8 * i) % 256).astype(np.uint8)
        rgb[disc] = (200, 60, 30)
    d = 255.0 * (0.6 * (0.5 + 0.5 * np.sin(2 * np.pi * x / w)) + 0.4 * y / h)
    d = np.clip(d, 0, 255).astype(np.uint8)
    d[disc] = 230
    depth = np.repeat(d[..., None], 3, axis=2)
    return np.ascontiguousarray(rgb), np.ascontiguousarray(depth)
