"""oracle/dibr.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy restatement, in float32 with one rounding per reference torch op, of the
DIBR (depth-image-based rendering) stage of VisionDepth3D:
    core/render_3d.py:135-143   frame_to_tensor / depth_to_tensor
    core/render_3d.py:145-172   estimate_subject_depth
    core/render_3d.py:175-187   enhance_curvature
    core/render_3d.py:198-216   suppress_artifacts_with_edge_mask
    core/render_3d.py:220-285   TemporalDepthFilter / DepthPercentileEMA / ConvergenceEMA
    core/render_3d.py:289-291   tensor_to_frame
    core/render_3d.py:328-374   feather_shift_edges
    core/render_3d.py:412-427   compute_dynamic_parallax_scale
    core/render_3d.py:463-511   ShiftSmoother / FloatingWindowTracker / FloatingBarEaser
    core/render_3d.py:519-558   shape_depth_for_pop
    core/render_3d.py:561-712   pixel_shift_cuda
    core/render_3d.py:717-892   sharpen / colour grade / DOF / format / anaglyph / side mask
    core/render_3d.py:895-929   FocalDepthTracker / compute_motion_metric
    core/render_3d.py:1086-1138,1227-1419  render_sbs_3d sizing + one loop iteration

Library ops the reference calls (torch 2.11 F.interpolate / F.grid_sample /
F.avg_pool2d / torch.quantile / torch.histc / torch.median, torchvision 0.26
gaussian_blur, cv2 4.13 cvtColor / filter2D / resize INTER_AREA) are restated
explicitly from their published definitions; tests/test_oracle_golden.py pins
each against the real op and against reference outputs in tests/golden/.
The reference's own last-ulp behaviour is machine dependent (torch.linspace
vector width, FMA contraction), so pinning is "<= 2e-6 on floats, <= 1 LSB on
uint8 with a counted number of flips", not bit-for-bit.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

f32 = np.float32


def _fma(a, b, c):
    """fp32 fused multiply-add (the fp64 product of two fp32 is exact; the fp64
    sum is exact or correctly rounded far below fp32 precision)."""
    return (np.asarray(a, dtype=np.float64) * np.asarray(b, dtype=np.float64)
            + np.asarray(c, dtype=np.float64)).astype(f32)


# --------------------------------------------------------------------------
# small library-op restatements
# --------------------------------------------------------------------------
def linspace32(n: int, start=-1.0, end=1.0) -> np.ndarray:
    """torch.linspace(start, end, n) float32 as torch 2.11 computes it on CPU and
    CUDA: step = (end-start)/(n-1) in fp32, then per element with a fused
    multiply-add:  i < n//2 : fma(step, i, start) ; else fma(-step, n-1-i, end).
    (The fp64 product/sum below is exact, so one rounding == fma.)"""
    if n == 1:
        return np.array([start], dtype=f32)
    start = f32(start)
    end = f32(end)
    step = np.float64(f32((end - start) / f32(n - 1)))
    i = np.arange(n).astype(np.float64)
    lo = (np.float64(start) + step * i).astype(f32)
    hi = (np.float64(end) - step * (n - 1 - i)).astype(f32)
    return np.where(np.arange(n) < n // 2, lo, hi).astype(f32)


def bgr_to_rgb01(frame_bgr: np.ndarray) -> np.ndarray:
    """frame_to_tensor (core/render_3d.py:135-138): u8 BGR HWC -> f32 RGB CHW /255."""
    rgb = frame_bgr[..., ::-1].astype(f32)
    return np.ascontiguousarray((rgb / f32(255.0)).astype(f32).transpose(2, 0, 1))


def bgr_to_gray_u8(frame_bgr: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(BGR2GRAY) for u8: (B*3735 + G*19235 + R*9798 + 2^14) >> 15."""
    b = frame_bgr[..., 0].astype(np.int64)
    g = frame_bgr[..., 1].astype(np.int64)
    r = frame_bgr[..., 2].astype(np.int64)
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def depth_bgr_to_01(depth_bgr: np.ndarray) -> np.ndarray:
    """depth_to_tensor (core/render_3d.py:140-143): gray u8 /255 -> f32 [1,H,W]."""
    return (bgr_to_gray_u8(depth_bgr).astype(f32) / f32(255.0)).astype(f32)[None]


def rgb01_to_bgr8(t: np.ndarray) -> np.ndarray:
    """tensor_to_frame (core/render_3d.py:289-291): *255, truncate, RGB->BGR."""
    v = (t.transpose(1, 2, 0).astype(f32) * f32(255)).astype(f32)
    return np.ascontiguousarray(v.astype(np.uint8)[..., ::-1])


def _resize_idx(insz: int, outsz: int):
    scale = f32(f32(insz) / f32(outsz))
    d = np.arange(outsz).astype(f32)
    s = _fma(scale, (d + f32(0.5)).astype(f32), f32(-0.5))  # torch contracts scale*(d+.5)-.5
    s = np.maximum(s, f32(0))
    i0 = s.astype(np.int64)
    i1 = i0 + (i0 < insz - 1)
    l1 = (s - i0.astype(f32)).astype(f32)
    l0 = (f32(1) - l1).astype(f32)
    return i0, i1, l0, l1


def resize_bilinear(src: np.ndarray, oh: int, ow: int) -> np.ndarray:
    """F.interpolate(mode='bilinear', align_corners=False) on [C,h,w] float32."""
    c, h, w = src.shape
    y0, y1, ly0, ly1 = _resize_idx(h, oh)
    x0, x1, lx0, lx1 = _resize_idx(w, ow)

    def hrow(r):  # fma(a, l0, b*l1) as the torch 2.11 CPU kernel rounds it
        b = (src[:, r][:, :, x1] * lx1).astype(f32)
        return _fma(src[:, r][:, :, x0], lx0, b)

    bot = (hrow(y1) * ly1[None, :, None]).astype(f32)
    return _fma(hrow(y0), ly0[None, :, None], bot)


def quantile32(arr: np.ndarray, q: float) -> np.float32:
    """torch.quantile(x, q) (linear): rank = fp32(q)*(n-1) in fp32; lerp as torch.lerp."""
    s = np.sort(arr.reshape(-1))
    n = s.size
    rank = f32(f32(q) * f32(n - 1))
    lo = int(np.floor(rank))
    hi = int(np.ceil(rank))
    w = f32(rank - f32(lo))
    a, b = s[lo], s[hi]
    diff = f32(b - a)
    if abs(w) < 0.5:
        return f32(a + f32(w * diff))
    return f32(b - f32(diff * f32(f32(1) - w)))


def avg_pool_same(x: np.ndarray, k: int) -> np.ndarray:
    """F.avg_pool2d(x[None], k, stride=1, padding=k//2)[0] on [C,H,W]; zero padding,
    count_include_pad=True (always / k*k), row-major fp32 accumulation.  Even k
    yields (H+1, W+1) like torch."""
    c, h, w = x.shape
    p = k // 2
    oh = h + 2 * p - k + 1
    ow = w + 2 * p - k + 1
    xp = np.zeros((c, h + 2 * p, w + 2 * p), dtype=f32)
    xp[:, p:p + h, p:p + w] = x
    acc = np.zeros((c, oh, ow), dtype=f32)
    for dy in range(k):
        for dx in range(k):
            acc = (acc + xp[:, dy:dy + oh, dx:dx + ow]).astype(f32)
    return (acc / f32(k * k)).astype(f32)


def _grad_mag(d: np.ndarray, absval: bool) -> np.ndarray:
    """sqrt(dx^2+dy^2) with backward differences, zero at x=0 / y=0 ([1,H,W])."""
    dx = np.zeros_like(d)
    dy = np.zeros_like(d)
    dx[:, :, 1:] = (d[:, :, 1:] - d[:, :, :-1]).astype(f32)
    dy[:, 1:, :] = (d[:, 1:, :] - d[:, :-1, :]).astype(f32)
    if absval:
        dx = np.abs(dx)
        dy = np.abs(dy)
    return np.sqrt(((dx * dx).astype(f32) + (dy * dy).astype(f32)).astype(f32)).astype(f32)


def grid_sample_border(img: np.ndarray, gx: np.ndarray, gy: np.ndarray) -> np.ndarray:
    """F.grid_sample(bilinear, padding_mode='border', align_corners=True) on [C,H,W]."""
    c, h, w = img.shape
    ix = ((gx + f32(1)).astype(f32) * f32(f32(w - 1) / f32(2))).astype(f32)
    iy = ((gy + f32(1)).astype(f32) * f32(f32(h - 1) / f32(2))).astype(f32)
    ix = np.minimum(f32(w - 1), np.maximum(ix, f32(0)))
    iy = np.minimum(f32(h - 1), np.maximum(iy, f32(0)))
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    wx = (ix - x0).astype(f32)
    ex = (f32(1) - wx).astype(f32)
    ny = (iy - y0).astype(f32)
    sy = (f32(1) - ny).astype(f32)
    x0 = x0.astype(np.int64)
    y0 = y0.astype(np.int64)
    x1 = np.minimum(x0 + 1, w - 1)
    y1 = np.minimum(y0 + 1, h - 1)
    nw = (sy * ex).astype(f32)
    ne = (sy * wx).astype(f32)
    sw = (ny * ex).astype(f32)
    se = (ny * wx).astype(f32)
    # torch's CPU kernel accumulates with fused multiply-adds, NW first
    out = (img[:, y0, x0] * nw).astype(f32)
    out = _fma(img[:, y0, x1], ne, out)
    out = _fma(img[:, y1, x0], sw, out)
    out = _fma(img[:, y1, x1], se, out)
    return out


# --------------------------------------------------------------------------
# statistics
# --------------------------------------------------------------------------
def subject_depth(d: np.ndarray) -> np.float32:
    """estimate_subject_depth (core/render_3d.py:145-172) on [1,H,W]."""
    _, h, w = d.shape
    crop = d[:, h // 5:h * 4 // 5, w // 5:w * 4 // 5]
    valid = crop[(crop > f32(0.05)) & (crop < f32(0.95))]
    if valid.size < 20:
        return f32(0.5)
    bins = np.floor((valid * f32(64)).astype(f32)).astype(np.int64)
    bins = np.clip(bins, 0, 63)
    hist = np.bincount(bins, minlength=64)
    peak = int(np.argmax(hist))  # first maximum
    subj = f32(f32(f32(peak) + f32(0.5)) * f32(1.0 / 64))
    s = np.sort(valid)
    med = s[(s.size - 1) // 2]  # torch.median -> lower middle
    v = f32(f32(f32(0.7) * subj) + f32(f32(0.3) * med))
    return f32(min(max(v, f32(0)), f32(1)))


def dynamic_parallax_scale(d: np.ndarray, min_scale=0.6, max_scale=1.0) -> float:
    """compute_dynamic_parallax_scale (core/render_3d.py:412-427) -> python float."""
    _, h, w = d.shape
    c = d[:, h // 4:h * 3 // 4, w // 4:w * 3 // 4].astype(np.float64)
    n = c.size
    mean64 = c.sum() / n
    mean = f32(mean64)
    var = f32(((c - mean64) ** 2).sum() / (n - 1))
    nv = f32(var / f32(mean + f32(1e-5)))
    nv = f32(min(max(nv, f32(0)), f32(1)))
    scale = f32(f32(min_scale) + f32(nv * f32(max_scale - min_scale)))
    return float(scale)


def motion_metric(prev_d, curr_d) -> float:
    """compute_motion_metric (core/render_3d.py:924-929)."""
    if prev_d is None:
        return 0.0
    mad = float(f32(np.abs((curr_d - prev_d).astype(f32)).astype(np.float64).mean()))
    return max(0.0, min(1.0, mad * 4.0))


# --------------------------------------------------------------------------
# temporal state (core/render_3d.py:220-285, 463-511, 895-922)
# --------------------------------------------------------------------------
@dataclass
class GlobalState:
    """Module-level singletons of the reference (survive across renders)."""
    pct_lo: np.float32 | None = None          # depth_ema_norm._lo (284)
    pct_hi: np.float32 | None = None
    conv_val: float | None = None             # conv_ema.val (285)
    fw_prev: float = 0.0                      # floating_window_tracker (500)
    fw_count: int = 0
    bar_prev: int = 0                         # bar_easer (511)


@dataclass
class ClipState:
    """Per-render objects created at core/render_3d.py:1174-1182."""
    tdf_prev: np.ndarray | None = None        # TemporalDepthFilter(alpha=0.5)
    sm_fg: float | None = None                # ShiftSmoother(0.15)
    sm_mg: float | None = None
    sm_bg: float | None = None
    focal: float | None = None                # FocalDepthTracker
    focal_alpha: float = 0.15
    prev_depth: np.ndarray | None = None      # prev_depth_tensor (1463)


def temporal_smooth(cs: ClipState, curr: np.ndarray, alpha=0.5) -> np.ndarray:
    if cs.tdf_prev is None:
        cs.tdf_prev = curr.copy()
    cs.tdf_prev = ((f32(alpha) * cs.tdf_prev).astype(f32)
                   + (f32(1 - alpha) * curr).astype(f32)).astype(f32)
    return cs.tdf_prev


def percentile_normalize(gs: GlobalState, depth01: np.ndarray,
                         p_lo=0.02, p_hi=0.98, alpha=0.92) -> np.ndarray:
    """DepthPercentileEMA.normalize (core/render_3d.py:241-262)."""
    d = np.clip(depth01, f32(0), f32(1)).astype(f32)
    lo = quantile32(d, p_lo)
    hi = quantile32(d, p_hi)
    if f32(hi - lo) < f32(1e-5):
        return d
    if gs.pct_lo is None:
        gs.pct_lo, gs.pct_hi = lo, hi
    else:
        gs.pct_lo = f32(f32(f32(alpha) * gs.pct_lo) + f32(f32(1 - alpha) * lo))
        gs.pct_hi = f32(f32(f32(alpha) * gs.pct_hi) + f32(f32(1 - alpha) * hi))
    den = f32(f32(gs.pct_hi - gs.pct_lo) + f32(1e-6))
    out = ((d - gs.pct_lo).astype(f32) / den).astype(f32)
    return np.clip(out, f32(0), f32(1)).astype(f32)


def shift_smooth(cs: ClipState, fg, mg, bg, alpha=0.15):
    if cs.sm_fg is None:
        cs.sm_fg, cs.sm_mg, cs.sm_bg = fg, mg, bg
    else:
        cs.sm_fg = alpha * fg + (1 - alpha) * cs.sm_fg
        cs.sm_mg = alpha * mg + (1 - alpha) * cs.sm_mg
        cs.sm_bg = alpha * bg + (1 - alpha) * cs.sm_bg
    return cs.sm_fg, cs.sm_mg, cs.sm_bg


def floating_window_smooth(gs: GlobalState, cur: float, threshold=0.002, alpha=0.97) -> float:
    if abs(cur - gs.fw_prev) < threshold:
        return gs.fw_prev
    gs.fw_prev = alpha * gs.fw_prev + (1 - alpha) * cur
    gs.fw_count += 1
    if gs.fw_count >= 100:
        gs.fw_prev = max(min(gs.fw_prev, 1.0), -1.0)
        gs.fw_count = 0
    return gs.fw_prev


def focal_update(cs: ClipState, candidate: float, motion: float,
                 deadband=0.03, max_step=0.02) -> float:
    cs.focal_alpha = 0.10 + 0.20 * max(0.0, min(1.0, float(motion)))
    c = float(candidate)
    if cs.focal is None:
        cs.focal = c
        return cs.focal
    if abs(c - cs.focal) < deadband:
        c = cs.focal
    new = (1.0 - cs.focal_alpha) * cs.focal + cs.focal_alpha * c
    delta = new - cs.focal
    if delta > max_step:
        new = cs.focal + max_step
    elif delta < -max_step:
        new = cs.focal - max_step
    cs.focal = max(0.0, min(1.0, new))
    return cs.focal


# --------------------------------------------------------------------------
# pixel_shift_cuda (core/render_3d.py:561-712)
# --------------------------------------------------------------------------
@dataclass
class ShiftParams:
    fg_shift: float = 4.5
    mg_shift: float = -1.5
    bg_shift: float = -6.0
    blur_ksize: int = 9
    feather_strength: float = 10.0
    max_pixel_shift_percent: float = 0.02
    parallax_balance: float = 0.8
    zero_parallax_strength: float = 0.0
    use_subject_tracking: bool = True
    enable_floating_window: bool = True
    enable_feathering: bool = True
    enable_edge_masking: bool = True
    convergence_strength: float = 0.0
    enable_dynamic_convergence: bool = True
    depth_pop_gamma: float = 0.85
    depth_pop_mid: float = 0.50
    depth_stretch_lo: float = 0.05
    depth_stretch_hi: float = 0.95
    fg_pop_multiplier: float = 1.20
    bg_push_multiplier: float = 1.10
    subject_lock_strength: float = 1.00


def curvature(d: np.ndarray, strength=0.08) -> np.ndarray:
    """enhance_curvature (175-187): d + (1 - (x^2+y^2)) * strength."""
    _, h, w = d.shape
    yy = linspace32(h)[:, None]
    xx = linspace32(w)[None, :]
    r2 = ((xx * xx).astype(f32) + (yy * yy).astype(f32)).astype(f32)
    curv = (f32(1) - r2).astype(f32)
    return (d + (curv * f32(strength)).astype(f32)[None]).astype(f32)


def shape_for_pop(d01: np.ndarray, subj: np.float32, p: ShiftParams):
    """shape_depth_for_pop (519-558). Returns (shaped, lo, hi)."""
    d = np.clip(d01, f32(0), f32(1)).astype(f32)
    lo = quantile32(d, p.depth_stretch_lo)
    hi = quantile32(d, p.depth_stretch_hi)
    subj = f32(min(max(subj, f32(0)), f32(1)))
    if f32(hi - lo) < f32(1e-5):
        ds = d
        ss = subj
    else:
        den = f32(f32(hi - lo) + f32(1e-6))
        ds = np.clip(((d - lo).astype(f32) / den).astype(f32), f32(0), f32(1)).astype(f32)
        ss = f32(min(max(f32(f32(subj - lo) / den), f32(0)), f32(1)))
    mid = f32(p.depth_pop_mid)
    centered = ((ds - ss).astype(f32) + mid).astype(f32)
    x = (centered - mid).astype(f32)
    # |x|^gamma: the correctly rounded fp32 value (fp64 pow, one rounding); torch's Sleef powf
    # is within 1 ulp of it
    pw = np.power(np.abs(x).astype(np.float64), np.float64(f32(p.depth_pop_gamma))).astype(f32)
    shaped = ((np.sign(x) * pw).astype(f32) + mid).astype(f32)
    return np.clip(shaped, f32(0), f32(1)).astype(f32), lo, hi


def edge_suppress(d: np.ndarray, total_shift: np.ndarray, feather: float, thr=0.02):
    """suppress_artifacts_with_edge_mask (198-216)."""
    g = _grad_mag(d, absval=True)
    z = (((g - f32(thr)).astype(f32) * f32(feather)).astype(f32) * f32(5)).astype(f32)
    ez = np.exp((-z).astype(np.float64)).astype(f32)  # correctly rounded fp32 exp
    edge = (f32(1) / (f32(1) + ez).astype(f32)).astype(f32)
    smooth = avg_pool_same((f32(1) - edge).astype(f32), 5)
    return (total_shift * smooth).astype(f32)


def feather_edges(shifted: np.ndarray, original: np.ndarray, wdepth: np.ndarray,
                  k: int, feather: float) -> np.ndarray:
    """feather_shift_edges (328-374)."""
    g = _grad_mag(wdepth, absval=False)
    edge = np.clip((g * f32(feather)).astype(f32), f32(0), f32(1)).astype(f32)
    blur = avg_pool_same(edge, k)
    mh = min(shifted.shape[1], blur.shape[1])
    mw = min(shifted.shape[2], blur.shape[2])
    b = blur[:, :mh, :mw]
    s = shifted[:, :mh, :mw]
    o = original[:, :mh, :mw]
    out = ((s * (f32(1) - b).astype(f32)).astype(f32) + (o * b).astype(f32)).astype(f32)
    return np.clip(out, f32(0), f32(1)).astype(f32)


def pixel_shift(gs: GlobalState, frame_t: np.ndarray, depth_t: np.ndarray,
                width: int, height: int, p: ShiftParams, return_parts=False):
    """pixel_shift_cuda (561-712). frame_t [3,h,w] RGB 0..1, depth_t [1,h,w].
    Returns (left_bgr_u8, right_bgr_u8, final_shift[1,H,W]) (+ dict of intermediates)."""
    W, H = int(width), int(height)
    frame = resize_bilinear(frame_t, H, W)
    d = resize_bilinear(depth_t, H, W)
    d = np.clip(curvature(d, 0.08), f32(0), f32(1)).astype(f32)
    subj_raw = subject_depth(d)
    d_sh, lo, hi = shape_for_pop(d, subj_raw, p)
    subj = subject_depth(d_sh)

    omd = (f32(1) - d_sh).astype(f32).astype(np.float64)
    fgw = np.clip((omd * np.sqrt(omd)).astype(f32), f32(0), f32(1))  # (1-d)^1.5 correctly rounded
    mgw = np.clip((f32(1) - (np.abs((d_sh - f32(p.depth_pop_mid)).astype(f32)) * f32(3)).astype(f32)
                   ).astype(f32), f32(0), f32(1))
    bgw = np.clip(d_sh, f32(0), f32(1))
    half = W / 2.0
    raw = (((fgw * f32(p.fg_shift)).astype(f32) * f32(p.fg_pop_multiplier)).astype(f32)
           + (mgw * f32(p.mg_shift)).astype(f32)).astype(f32)
    raw = (raw + ((bgw * f32(p.bg_shift)).astype(f32) * f32(p.bg_push_multiplier)).astype(f32)
           ).astype(f32)
    total = ((raw * f32(p.parallax_balance)).astype(f32) / f32(half)).astype(f32)

    zpo = 0.0
    if p.use_subject_tracking:
        a = f32(subj * f32(p.parallax_balance))
        t1 = f32(f32(f32(-a) * f32(p.fg_shift)) * f32(p.fg_pop_multiplier))
        t2 = f32(f32(-a) * f32(p.mg_shift))
        t3 = f32(f32(a * f32(p.bg_shift)) * f32(p.bg_push_multiplier))
        z = f32(f32(f32(t1 + t2) + t3) / f32(half))
        z = f32(z * f32(float(p.subject_lock_strength)))
        z = f32(z - f32(float(p.zero_parallax_strength)))
        if p.enable_floating_window:
            sw = f32(min(max(f32(f32(1) - f32(subj * f32(2))), f32(0.5)), f32(1)))
            z = f32(z * sw)
            z = f32(min(max(z, f32(-0.35)), f32(0.35)))
            zpo = floating_window_smooth(gs, float(z), threshold=0.0015)
        else:
            zpo = float(z)
        total = (total - f32(zpo)).astype(f32)

    max_norm = (W * p.max_pixel_shift_percent) / half
    total = np.clip(total, f32(-max_norm), f32(max_norm)).astype(f32)

    if p.convergence_strength != 0.0:
        if p.enable_dynamic_convergence:
            conv = float(f32(subject_depth(d_sh) * f32(p.convergence_strength)))
        else:
            conv = p.convergence_strength
        total = (total - f32(conv / half)).astype(f32)

    mask_strength = float(np.clip(p.feather_strength / 10.0, 0.05, 0.3))
    if p.enable_edge_masking:
        sup = edge_suppress(d_sh, total, p.feather_strength)
        final = ((f32(1.0 - mask_strength) * total).astype(f32)
                 + (f32(mask_strength) * sup).astype(f32)).astype(f32)
    else:
        final = total

    xs = linspace32(W)[None, :]
    ys = np.broadcast_to(linspace32(H)[:, None], (H, W))
    sv = final[0]
    gxl = (xs + sv).astype(f32)
    gxr = (xs - sv).astype(f32)
    wl = grid_sample_border(frame, gxl, ys)
    wr = grid_sample_border(frame, gxr, ys)
    if p.enable_feathering:
        wdl = grid_sample_border(d_sh, gxl, ys)
        wdr = grid_sample_border(d_sh, gxr, ys)
        lb = feather_edges(wl, frame, wdl, p.blur_ksize, p.feather_strength)
        rb = feather_edges(wr, frame, wdr, p.blur_ksize, p.feather_strength)
    else:
        lb, rb = wl, wr
    left = rgb01_to_bgr8(lb)
    right = rgb01_to_bgr8(rb)
    if return_parts:
        return left, right, final, dict(subj_raw=subj_raw, lo=lo, hi=hi, subj=subj,
                                        d_shaped=d_sh, zpo=zpo, left_f=lb, right_f=rb)
    return left, right, final


def heal_missing_pixels(warped: np.ndarray, original: np.ndarray, edge_mask=None, heal_strength=0.5) -> np.ndarray:
    """heal_missing_pixels (core/render_3d.py:431-459; defined but never called by the reference --
    the "gradient-blend occlusion fill" of its method notes).  [3,H,W] float32 in, out."""
    gray = ((warped[0] + warped[1]).astype(f32) + warped[2]).astype(f32)
    gray = (gray / f32(3)).astype(f32)[None]
    g = _grad_mag(gray, absval=False)
    miss = (g > f32(0.05)).astype(f32)
    miss = np.clip(avg_pool_same(miss, 5), f32(0), f32(1)).astype(f32)
    if edge_mask is not None:
        miss = np.maximum(miss, edge_mask.astype(f32))
    hm = (f32(heal_strength) * miss).astype(f32)
    healed = (((f32(1) - hm).astype(f32) * warped).astype(f32) + (hm * original).astype(f32)).astype(f32)
    soft = avg_pool_same(healed, 3)
    sm = (f32(0.3) * miss).astype(f32)
    out = (((f32(1) - sm).astype(f32) * healed).astype(f32) + (sm * soft).astype(f32)).astype(f32)
    return np.clip(out, f32(0), f32(1)).astype(f32)


# --------------------------------------------------------------------------
# post chain (core/render_3d.py:717-892)
# --------------------------------------------------------------------------
def gaussian_kernel1d(ksize: int, sigma: float) -> np.ndarray:
    """torchvision _get_gaussian_kernel1d: pdf on linspace(-(k-1)/2,(k-1)/2,k), normalised."""
    half = (ksize - 1) * 0.5
    x = linspace32(ksize, -half, half)
    q = (x / f32(sigma)).astype(f32)
    pdf = np.exp((f32(-0.5) * (q * q).astype(f32)).astype(f32).astype(np.float64)).astype(f32)
    tot = f32(0)
    for v in pdf:  # sequential fp32 sum
        tot = f32(tot + v)
    return (pdf / tot).astype(f32)


def gaussian_blur(img: np.ndarray, ksize: int, sigma: float) -> np.ndarray:
    """torchvision gaussian_blur: reflect pad, depthwise conv with outer(k1d,k1d)."""
    k1 = gaussian_kernel1d(ksize, sigma)
    k2 = (k1[:, None] * k1[None, :]).astype(f32)
    p = ksize // 2
    xp = np.pad(img, ((0, 0), (p, p), (p, p)), mode="reflect")
    c, h, w = img.shape
    acc = np.zeros_like(img, dtype=f32)
    for dy in range(ksize):
        for dx in range(ksize):
            acc = _fma(xp[:, dy:dy + h, dx:dx + w], k2[dy, dx], acc)  # row-major fma chain
    return acc


def apply_dof(rgb: np.ndarray, depth: np.ndarray, focal: float, max_sigma=2.0,
              focus_width=0.35, num_levels=5) -> np.ndarray:
    """apply_dof_cuda (769-834)."""
    diff = np.abs((depth - f32(focal)).astype(f32))
    bw = np.clip((diff / f32(focus_width + 1e-6)).astype(f32), f32(0), f32(1))
    sig = linspace32(num_levels, 0.0, float(max_sigma))
    levels = []
    for s in sig:
        if float(s) == 0.0:
            levels.append(rgb)
        else:
            k = int(2 * math.ceil(2 * float(s)) + 1)
            levels.append(gaussian_blur(rgb, k, float(s)))
    stack = np.stack(levels, 0)
    n = num_levels
    bidx = np.clip((bw * f32(n - 1)).astype(f32), f32(0), f32(n - 1 - 1e-6))
    lo = np.clip(np.floor(bidx).astype(np.int64), 0, n - 2)
    a = (bidx - lo.astype(f32)).astype(f32)[0]
    lov = np.take_along_axis(stack, np.broadcast_to(lo[None], (1, 3) + lo.shape[1:]), 0)[0]
    upv = np.take_along_axis(stack, np.broadcast_to(lo[None] + 1, (1, 3) + lo.shape[1:]), 0)[0]
    out = (((f32(1) - a).astype(f32) * lov).astype(f32) + (a * upv).astype(f32)).astype(f32)
    return np.clip(out, f32(0), f32(1)).astype(f32)


def color_grade(rgb: np.ndarray, sat=1.0, con=1.0, bri=0.0) -> np.ndarray:
    """apply_color_grade (734-767)."""
    r, g, b = rgb[0], rgb[1], rgb[2]
    luma = (((f32(0.2126) * r).astype(f32) + (f32(0.7152) * g).astype(f32)).astype(f32)
            + (f32(0.0722) * b).astype(f32)).astype(f32)
    out = []
    for c in (r, g, b):
        s = (luma + ((c - luma).astype(f32) * f32(sat)).astype(f32)).astype(f32)
        s = (f32(0.5) + ((s - f32(0.5)).astype(f32) * f32(con)).astype(f32)).astype(f32)
        s = (s + f32(bri)).astype(f32)
        out.append(s)
    return np.clip(np.stack(out, 0), f32(0), f32(1)).astype(f32)


def _round_half_even_u8(v: np.ndarray) -> np.ndarray:
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def sharpen(frame: np.ndarray, factor=1.0) -> np.ndarray:
    """apply_sharpening (717-732): cv2.filter2D u8, 3x3 cross kernel normalised in
    fp32, BORDER_REFLECT_101, fp32 accumulation, round-half-even, saturate."""
    kern = np.array([[0, -1, 0], [-1, 5 + factor, -1], [0, -1, 0]], dtype=f32)
    ks = kern.sum(dtype=f32)
    if ks != 0:
        kern = (kern / ks).astype(f32)
    xp = np.pad(frame.astype(f32), ((1, 1), (1, 1), (0, 0)), mode="reflect")
    h, w = frame.shape[:2]
    acc = np.zeros(frame.shape, dtype=f32)
    for (dy, dx) in ((0, 1), (1, 0), (1, 1), (1, 2), (2, 1)):
        acc = _fma(xp[dy:dy + h, dx:dx + w], kern[dy, dx], acc)  # cv2: row-major fma chain
    return _round_half_even_u8(acc)


def resize_area_int(img: np.ndarray, ow: int, oh: int) -> np.ndarray:
    """cv2.resize(INTER_AREA) for integer shrink factors (and identity):
    box sum * fp32(1/area), round-half-even.  2x2 uses the integer fast path
    (sum + 2) >> 2."""
    h, w = img.shape[:2]
    if (w, h) == (ow, oh):
        return img.copy()
    assert w % ow == 0 and h % oh == 0, "integer INTER_AREA factors only (resize_area handles the rest)"
    sx, sy = w // ow, h // oh
    v = img.reshape(oh, sy, ow, sx, -1).astype(np.int64).sum(axis=(1, 3))
    if sx == 2 and sy == 2:
        return ((v + 2) >> 2).astype(np.uint8)
    return _round_half_even_u8((v.astype(f32) * f32(1.0 / (sx * sy))).astype(f32))


def area_tab(ssize: int, dsize: int):
    """cv2's computeResizeAreaTab (imgproc resize.cpp; third-party, opencv 4.13 as installed): per destination
    index the list of (source index, fp32 weight).  All geometry in double, weights rounded once to fp32."""
    import math
    scale = 1.0 / (dsize / ssize)  # cv2: inv_scale = dsize / ssize; scale = 1. / inv_scale
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1 = math.ceil(fsx1)
        sx2 = min(math.floor(fsx2), ssize - 1)
        sx1 = min(sx1, sx2)
        ent = []
        if sx1 - fsx1 > 1e-3:
            ent.append((sx1 - 1, f32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            ent.append((sx, f32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            ent.append((sx2, f32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
        tab.append(ent)
    return tab


def _area_linear_tab(ssize: int, dsize: int):
    """cv2 resize(): coefficient table of the bilinear scheme INTER_AREA falls back to when an axis is enlarged
    ("area_mode" branch of the generic path): source index and two 11-bit fixed-point weights per destination."""
    import math
    scale = 1.0 / (dsize / ssize)
    inv = dsize / ssize
    ofs = np.zeros(dsize, dtype=np.int64)
    a = np.zeros((dsize, 2), dtype=np.int64)
    for dx in range(dsize):
        sx = math.floor(dx * scale)
        fx = f32((dx + 1) - (sx + 1) * inv)
        fx = f32(0) if fx <= 0 else f32(fx - math.floor(fx))
        if sx < 0:
            fx, sx = f32(0), 0
        if sx >= ssize - 1:
            fx, sx = f32(0), ssize - 1
        a[dx, 0] = int(np.rint(f32(f32(f32(1.0) - fx) * f32(2048))))  # saturate_cast<short>(c * INTER_RESIZE_COEF_SCALE)
        a[dx, 1] = int(np.rint(f32(fx * f32(2048))))
        ofs[dx] = sx
    return ofs, a


def _resize_area_enlarge(img: np.ndarray, ow: int, oh: int) -> np.ndarray:
    """cv2.resize(INTER_AREA) when at least one axis grows: fixed-point bilinear with area-style coefficients
    (HResizeLinear int rows, VResizeLinear<uchar,int,short>: ((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2)."""
    h, w = img.shape[:2]
    xo, xa = _area_linear_tab(w, ow)
    yo, ya = _area_linear_tab(h, oh)
    S = img.astype(np.int64)
    x1 = np.minimum(xo + 1, w - 1)
    rows = S[:, xo, :] * xa[None, :, 0, None] + S[:, x1, :] * xa[None, :, 1, None]
    y1 = np.minimum(yo + 1, h - 1)
    b0, b1 = ya[:, 0][:, None, None], ya[:, 1][:, None, None]
    out = (((b0 * (rows[yo] >> 4)) >> 16) + ((b1 * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def resize_area(img: np.ndarray, ow: int, oh: int) -> np.ndarray:
    """cv2.resize(img, (ow, oh), interpolation=cv2.INTER_AREA) as called by pad_to_aspect_ratio / the Half-SBS eye fit
    (core/render_3d.py:121, 1413-1414).  Identity -> copy; integer factors on both axes -> cv2's "area fast" path
    (resize_area_int); both axes shrinking -> cv2's ResizeArea_: per source row buf[dx] = sum_k S[sx_k] * alpha_k
    (fp32, separate multiply and add, table order), rows combined as sum = beta_0 * buf_0 (+ beta_j * buf_j ...) in
    fp32, saturate_cast<uchar> = round-half-even; any axis enlarging -> cv2's fixed-point bilinear emulation
    (_resize_area_enlarge; e.g. 1280x720 eyes into the hard-coded 1920x1080 Full-SBS canvas, core/render_3d.py:1121).
    Every branch is pinned against the real cv2.resize in tests/test_oracle_golden.py (exact)."""
    h, w = img.shape[:2]
    if (w, h) == (ow, oh):
        return img.copy()
    if w % ow == 0 and h % oh == 0:
        return resize_area_int(img, ow, oh)
    if ow > w or oh > h:
        assert img.ndim == 3, "enlarging branch restated for BGR images"
        return _resize_area_enlarge(img, ow, oh)
    squeeze = img.ndim == 2
    S = (img[..., None] if squeeze else img).astype(f32)
    xt, yt = area_tab(w, ow), area_tab(h, oh)
    buf = np.zeros((h, ow, S.shape[2]), dtype=f32)
    for k in range(max(len(e) for e in xt)):
        idx = np.array([e[k][0] if k < len(e) else 0 for e in xt])
        al = np.array([e[k][1] if k < len(e) else f32(0) for e in xt], dtype=f32)
        valid = np.array([k < len(e) for e in xt])
        term = (S[:, idx, :] * al[None, :, None]).astype(f32)
        buf = np.where(valid[None, :, None], (buf + term).astype(f32), buf)
    out = np.zeros((oh, ow, S.shape[2]), dtype=np.uint8)
    for dy, ent in enumerate(yt):
        acc = None
        for (sy, beta) in ent:
            t = (beta * buf[sy]).astype(f32)
            acc = t if acc is None else (acc + t).astype(f32)
        out[dy] = _round_half_even_u8(acc)
    return out[..., 0] if squeeze else out


def pad_to_aspect(img: np.ndarray, tw: int, th: int) -> np.ndarray:
    """pad_to_aspect_ratio (101-131)."""
    h, w = img.shape[:2]
    ta = tw / th
    ca = w / h
    if ca > ta:
        nw, nh = tw, int(tw / ca)
    else:
        nh, nw = th, int(ca * th)
    rs = resize_area(img, nw, nh)
    out = np.zeros((th, tw, 3), dtype=np.uint8)
    xo = (tw - nw) // 2
    yo = (th - nh) // 2
    out[yo:yo + nh, xo:xo + nw] = rs
    return out


def anaglyph(left: np.ndarray, right: np.ndarray) -> np.ndarray:
    """generate_anaglyph_3d (862-883) on channel indices 0,1,2 as written."""
    l = (left.astype(f32) / f32(255)).astype(f32)
    r = (right.astype(f32) / f32(255)).astype(f32)

    def mix(img, c0, c1, c2):
        a = (f32(c0) * img[..., 0]).astype(f32)
        b = (f32(abs(c1)) * img[..., 1]).astype(f32)
        a = (a + b).astype(f32) if c1 >= 0 else (a - b).astype(f32)
        c = (f32(abs(c2)) * img[..., 2]).astype(f32)
        return (a + c).astype(f32) if c2 >= 0 else (a - c).astype(f32)

    red = mix(l, 0.4561, 0.5005, 0.1762)
    grn = mix(r, 0.3764, 0.7616, -0.1876)
    blu = mix(r, -0.0401, -0.1126, 1.2723)
    out = np.stack([np.clip(red, 0, 1), np.clip(grn, 0, 1), np.clip(blu, 0, 1)], -1).astype(f32)
    return (out * f32(255)).astype(f32).astype(np.uint8)


def format_output(left: np.ndarray, right: np.ndarray, fmt: str) -> np.ndarray:
    """format_3d_output (837-860).  VR: the eyes arrive from pad_to_aspect_ratio already 1440 x 1600 (1129-1133,
    1415-1417), so format_3d_output's cv2.resize(..., (1440, 1600)) is the identity copy."""
    if fmt == "VR":
        assert left.shape[:2] == (1600, 1440) and right.shape[:2] == (1600, 1440), "VR eyes must be 1440x1600"
    if fmt == "Red-Cyan Anaglyph":
        return anaglyph(left, right)
    if fmt == "Passive Interlaced":
        out = np.zeros_like(left)
        out[::2] = left[::2]
        out[1::2] = right[1::2]
        return out
    return np.hstack((left, right))


def side_mask(img: np.ndarray, side: str, width: int) -> np.ndarray:
    """apply_side_mask (885-892)."""
    out = img.copy()
    w = img.shape[1]
    if side == "left":
        out[:, :width] = 0
    elif side == "right":
        out[:, w - width:] = 0
    return out


# --------------------------------------------------------------------------
# render_sbs_3d: sizing (1074-1138, 1250-1259) and one loop iteration (1227-1419)
# --------------------------------------------------------------------------
@dataclass
class RenderParams:
    output_width: int = 1920
    output_height: int = 1080
    fg_shift: float = 4.5
    mg_shift: float = -1.5
    bg_shift: float = -6.0
    sharpness_factor: float = 0.2
    output_format: str = "Half-SBS"
    aspect_ratio: float = 16 / 9
    dof_strength: float = 0.0
    feather_strength: float = 0.0
    blur_ksize: int = 1
    use_subject_tracking: bool = False
    use_floating_window: bool = False
    max_pixel_shift_percent: float = 0.02
    preserve_original_aspect: bool = False
    zero_parallax_strength: float = 0.0
    enable_edge_masking: bool = True
    enable_feathering: bool = True
    original_video_width: int | None = None
    original_video_height: int | None = None
    convergence_strength: float = 0.0
    enable_dynamic_convergence: bool = True
    ipd_factor: float = 1.0
    color_saturation: float = 1.0
    color_contrast: float = 1.0
    color_brightness: float = 0.0


@dataclass
class SizePlan:
    crop_x0: int
    crop_y0: int
    crop_w: int
    crop_h: int
    target_eye_w: int
    target_eye_h: int
    resized_width: int
    resized_height: int
    per_eye_w: int
    per_eye_h: int
    out_width: int
    out_height: int


def plan_sizes(src_w: int, src_h: int, rp: RenderParams) -> SizePlan:
    target_ratio = rp.aspect_ratio
    cw, ch, cx0, cy0 = src_w, src_h, 0, 0
    cur = src_w / src_h
    if abs(cur - target_ratio) > 0.01:
        if cur > target_ratio:
            cw = int(src_h * target_ratio)
            cx0 = (src_w - cw) // 2
        else:
            ch = int(src_w / target_ratio)
            cy0 = (src_h - ch) // 2
    fmt = rp.output_format
    if rp.preserve_original_aspect:
        ow = rp.original_video_width if rp.original_video_width is not None else src_w
        oh = rp.original_video_height if rp.original_video_height is not None else src_h
        if rp.original_video_width is None or rp.original_video_height is None:
            ow, oh = src_w, src_h
        rw, rh = ow, oh
        if fmt == "Full-SBS":
            pw, ph, outw, outh = rw, rh, rw * 2, rh
        elif fmt == "Half-SBS":
            pw, ph, outw, outh = rw // 2, rh, rw, rh
        elif fmt == "VR":
            pw, ph, outw, outh = 1440, 1600, 2880, 1600
        else:
            pw, ph, outw, outh = rw, rh, rw * 2, rh
        tew, teh = pw, ph
    else:
        rh = rp.output_height
        rw = int(rh * target_ratio)
        if rw % 2 != 0:
            rw += 1
        if fmt == "Full-SBS":
            pw, ph, outw, outh = 1920, 1080, 3840, 1080
        elif fmt == "Half-SBS":
            pw, ph, outw, outh = rw // 2, rh, rw, rh
        elif fmt == "VR":
            pw, ph, outw, outh = 1440, 1600, 2880, 1600
        else:
            pw, ph, outw, outh = rw, rh, rw * 2, rh
        tew = pw
        teh = int(pw / target_ratio)
        if teh % 2 != 0:
            teh += 1
    return SizePlan(cx0, cy0, cw, ch, tew, teh, rw, rh, pw, ph, outw, outh)


def render_frame(gs: GlobalState, cs: ClipState, frame_bgr: np.ndarray, depth_bgr: np.ndarray,
                 rp: RenderParams, return_parts=False):
    """One iteration of the render_sbs_3d loop body (core/render_3d.py:1227-1419),
    blank-frame bypass and auto-crop excluded (both default off)."""
    sh, sw = frame_bgr.shape[:2]
    pl = plan_sizes(sw, sh, rp)
    ft = bgr_to_rgb01(frame_bgr)
    dt = depth_bgr_to_01(depth_bgr)
    ft = ft[:, pl.crop_y0:pl.crop_y0 + pl.crop_h, pl.crop_x0:pl.crop_x0 + pl.crop_w]
    dt = dt[:, pl.crop_y0:pl.crop_y0 + pl.crop_h, pl.crop_x0:pl.crop_x0 + pl.crop_w]
    ft = resize_bilinear(ft, pl.target_eye_h, pl.target_eye_w)
    dt = resize_bilinear(dt, pl.target_eye_h, pl.target_eye_w)
    dt = temporal_smooth(cs, dt, 0.5)
    dn = percentile_normalize(gs, dt)
    fg, mg, bg = shift_smooth(cs, rp.fg_shift, rp.mg_shift, rp.bg_shift)
    dyn = dynamic_parallax_scale(dn, 0.90, 1.15)
    fg *= dyn
    mg *= dyn
    bg *= dyn
    if rp.ipd_factor != 0.0:
        fg *= rp.ipd_factor
        mg *= rp.ipd_factor
        bg *= rp.ipd_factor
    sp = ShiftParams(
        fg_shift=fg, mg_shift=mg, bg_shift=bg, blur_ksize=rp.blur_ksize,
        feather_strength=rp.feather_strength,
        max_pixel_shift_percent=rp.max_pixel_shift_percent,
        parallax_balance=0.8,  # never forwarded
        zero_parallax_strength=rp.zero_parallax_strength,
        use_subject_tracking=rp.use_subject_tracking,
        enable_floating_window=rp.use_floating_window,
        enable_feathering=rp.enable_feathering, enable_edge_masking=rp.enable_edge_masking,
        convergence_strength=rp.convergence_strength,
        enable_dynamic_convergence=rp.enable_dynamic_convergence)
    left, right, _ = pixel_shift(gs, ft, dn, pl.resized_width, pl.resized_height, sp)

    cand = subject_depth(dn)
    mot = motion_metric(cs.prev_depth, dn)
    focal = focal_update(cs, float(cand), mot)

    lt = bgr_to_rgb01(left)
    rt = bgr_to_rgb01(right)
    if rp.dof_strength > 0.0:
        hh, ww = lt.shape[1], lt.shape[2]
        dd = resize_bilinear(dn, hh, ww)
        lt = apply_dof(lt, dd, focal, max_sigma=rp.dof_strength, focus_width=0.35)
        rt = apply_dof(rt, dd, focal, max_sigma=rp.dof_strength, focus_width=0.35)
    lt = color_grade(lt, rp.color_saturation, rp.color_contrast, rp.color_brightness)
    rt = color_grade(rt, rp.color_saturation, rp.color_contrast, rp.color_brightness)
    left = rgb01_to_bgr8(lt)
    right = rgb01_to_bgr8(rt)

    sd = cand  # estimate_subject_depth(depth_tensor) again (1390): same input, same value
    half = f32(pl.resized_width / 2 + 1e-6)
    raw_zero = f32(f32(f32(f32(-sd) * f32(fg)) + f32(f32(-sd) * f32(mg))) + f32(sd * f32(bg)))
    raw_zero = float(f32(raw_zero / half))
    gs.conv_val = raw_zero if gs.conv_val is None else (0.97 * gs.conv_val + (1 - 0.97) * raw_zero)
    stable = gs.conv_val
    bar = 0
    if rp.use_floating_window and rp.use_subject_tracking:
        raw_bar = int(abs(stable) * pl.resized_width * 0.75)
        gs.bar_prev = int(0.85 * gs.bar_prev + (1 - 0.85) * raw_bar)
        bar = max(min(gs.bar_prev, 80), 0)
        if stable > 0.005:
            left = side_mask(left, "right", bar)
            right = side_mask(right, "right", bar)
        elif stable < -0.005:
            left = side_mask(left, "left", bar)
            right = side_mask(right, "left", bar)

    ls = sharpen(left, rp.sharpness_factor)
    rs = sharpen(right, rp.sharpness_factor)
    if rp.output_format == "Half-SBS":
        lo_ = resize_area(ls, pl.per_eye_w, pl.per_eye_h)
        ro_ = resize_area(rs, pl.per_eye_w, pl.per_eye_h)
    else:
        lo_ = pad_to_aspect(ls, pl.per_eye_w, pl.per_eye_h)
        ro_ = pad_to_aspect(rs, pl.per_eye_w, pl.per_eye_h)
    final = format_output(lo_, ro_, rp.output_format)
    cs.prev_depth = dn
    if return_parts:
        return final, dict(left=left, right=right, focal=focal, dyn=dyn, stable_zero=stable,
                           bar=bar, depth_norm=dn)
    return final


def _cubic_axis(ssize, dsize):
    """cv2.resize(INTER_CUBIC) source taps and Catmull-Rom-like weights (A = -0.75) of one axis, float32 like cv2's
    interpolateCubic (imgproc/src/resize.cpp); border taps are clamped."""
    scale = ssize / dsize
    d = np.arange(dsize)
    fx = ((d + 0.5) * scale - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int64)
    x = (fx - sx.astype(np.float32)).astype(np.float32)
    A, one = np.float32(-0.75), np.float32(1)
    x1, xm = x + one, one - x
    c0 = ((A * x1 - np.float32(5) * A) * x1 + np.float32(8) * A) * x1 - np.float32(4) * A
    c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    c2 = ((A + np.float32(2)) * xm - (A + np.float32(3))) * xm * xm + one
    c3 = one - c0 - c1 - c2
    ofs = np.clip(sx[:, None] - 1 + np.arange(4)[None], 0, ssize - 1)
    return ofs, np.stack([c0, c1, c2, c3], 1).astype(np.float32)


def resize_cubic_u8(img: np.ndarray, ow: int, oh: int) -> np.ndarray:
    """cv2.resize(u8 [h, w], (ow, oh), interpolation=cv2.INTER_CUBIC) as the depth writer applies it
    (core/render_depth.py:1917, 193).  The installed cv2 (4.13 with Intel IPP) evaluates the bicubic in float32:
    rows first, then columns, round half to even.  Pinned against cv2 at <= 1 LSB with < 1e-4 of the pixels off
    (tests/test_oracle_golden.py); cv2's own non-IPP fixed-point path differs from its IPP path by 1 LSB on 3-5 %."""
    h, w = img.shape
    if (h, w) == (oh, ow):
        return img.copy()
    xo, xa = _cubic_axis(w, ow)
    yo, ya = _cubic_axis(h, oh)
    S = img.astype(np.float32)
    R = S[yo]                                   # [oh, 4, w]
    V = ((R[:, 0] * ya[:, 0, None] + R[:, 1] * ya[:, 1, None]) + R[:, 2] * ya[:, 2, None]) + R[:, 3] * ya[:, 3, None]
    C = V[:, xo]                                # [oh, ow, 4]
    Hh = ((C[..., 0] * xa[None, :, 0] + C[..., 1] * xa[None, :, 1]) + C[..., 2] * xa[None, :, 2]) + C[..., 3] * xa[None, :, 3]
    return np.clip(np.rint(Hh), 0, 255).astype(np.uint8)
