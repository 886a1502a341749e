"""Deterministic synthetic clips; numpy only.

smooth set: ramps + moving disc (parity gate);  noise set: white-noise RGB
(throughput / adversarial LSB case).  Frame i uses default_rng(1000+i).
natural set: band-limited random colour / depth fields with fine texture, translating 3 px per frame (temporally
coherent like footage).  Sub-pixel warps of such content land on generic values, not on the k/255 truncation
boundaries the ramps of the smooth set sit on, so it is the set the 1-LSB gate against the reference is meant for.
"""
import numpy as np


def _field(rng, h, w, cell, ch):
    """Bilinear interpolation (float64, numpy only) of a coarse random grid: a band-limited field in [0, 1]."""
    gh, gw = h // cell + 2, w // cell + 2
    g = rng.random((gh, gw, ch))
    yy = np.arange(h, dtype=np.float64) / cell
    xx = np.arange(w, dtype=np.float64) / cell
    y0, x0 = yy.astype(np.int64), xx.astype(np.int64)
    fy, fx = (yy - y0)[:, None, None], (xx - x0)[None, :, None]
    a = g[y0][:, x0] * (1 - fx) + g[y0][:, x0 + 1] * fx
    b = g[y0 + 1][:, x0] * (1 - fx) + g[y0 + 1][:, x0 + 1] * fx
    return a * (1 - fy) + b * fy


def _natural(i, w, h):
    rng = np.random.default_rng(777)           # one scene; frame i is a translated view of it
    pad = 3 * 64                               # room for the 3 px / frame pan
    big_w = w + pad
    col = 0.15 + 0.7 * _field(rng, h, big_w, max(8, h // 12), 3)
    col += 0.10 * (_field(rng, h, big_w, 3, 3) - 0.5)       # fine texture
    dep = 0.1 + 0.8 * _field(rng, h, big_w, max(12, h // 6), 1)[..., 0]
    dep += 0.04 * (_field(rng, h, big_w, 5, 1)[..., 0] - 0.5)
    off = (3 * i) % pad
    rgb = np.clip(col[:, off:off + w] * 255.0, 0, 255).astype(np.uint8)
    d = np.clip(dep[:, off:off + w] * 255.0, 0, 255).astype(np.uint8)
    return rgb, d


def synth_frame(i: int, w: int, h: int, kind: str = "smooth"):
    """Return (frame_bgr u8 [h,w,3], depth_bgr u8 [h,w,3]) for frame index i."""
    rng = np.random.default_rng(1000 + i)
    y, x = np.mgrid[0:h, 0:w]
    cx = (w // 4 + 20 * i) % w
    cy = h // 2
    disc = (x - cx) ** 2 + (y - cy) ** 2 <= (h // 6) ** 2
    if kind == "natural":
        rgb, d = _natural(i, w, h)
        d = d.copy()
        d[disc] = 230
        rgb = rgb.copy()
        rgb[disc] = (rgb[disc].astype(np.int32) * 2 // 3 + 60).astype(np.uint8)
        return np.ascontiguousarray(rgb), np.ascontiguousarray(np.repeat(d[..., None], 3, axis=2))
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
