"""Drop-in for the hot-path surface of the reference's core/render_3d.py, backed by
libvd3d.so (hand-written sm_100a CUDA).  Same names, argument meaning, return types and
error behaviour as the reference (file:line cited per function); no Tk, no ONNXRuntime,
no CPU fallback.

    from visiondepth3d_b200 import render_3d as core_render_3d
    left, right, shift = core_render_3d.pixel_shift_cuda(frame_t, depth_t, W, H, 4.5, -1.5, -6.0)
"""
import ctypes as C
import threading
import time

import numpy as np

from . import _lib
from ._lib import FrameInfo, RenderParams, ShiftParams, SizePlan

try:  # torch is plumbing only (device tensors in / out); numpy works too
    import torch
except Exception:  # pragma: no cover
    torch = None

# module-level names the reference's callers import (core/render_3d.py:33-47)
suspend_flag = threading.Event()
cancel_flag = threading.Event()
aspect_ratios = {
    "Default (16:9)": 16 / 9,
    "CinemaScope (2.39:1)": 2.39,
    "21:9 UltraWide": 21 / 9,
    "4:3 (Classic Films)": 4 / 3,
    "1:1 (Square)": 1 / 1,
    "2.35:1 (Classic Cinematic)": 2.35,
    "2.76:1 (Ultra-Panavision)": 2.76,
}

torch_device = None
if torch is not None:
    torch_device = torch.device("cuda" if torch.cuda.is_available() else "cpu")


def _ctx():
    dev = 0
    if torch is not None and torch.cuda.is_available():
        dev = torch.cuda.current_device()
    return _lib.default_context(dev)


def reset_temporal_state(global_state=True, clip_state=True):
    """Reset what the reference keeps in module singletons (core/render_3d.py:284-285,500,511)
    and in per-render objects (1174-1182)."""
    which = (_lib.STATE_GLOBAL if global_state else 0) | (_lib.STATE_CLIP if clip_state else 0)
    _ctx().reset(which)


# ---------------------------------------------------------------------------
# converters (core/render_3d.py:135-143, 289-291) -- host-side glue the reference's callers use
# ---------------------------------------------------------------------------
def frame_to_tensor(frame):
    """BGR u8 HWC -> RGB f32 CHW /255 on torch_device (core/render_3d.py:135-138)."""
    t = torch.from_numpy(np.ascontiguousarray(frame[..., ::-1])).to(torch_device)
    return t.float().permute(2, 0, 1) / 255.0


def depth_to_tensor(depth_frame):
    """cv2 BGR2GRAY (fixed point) then /255 (core/render_3d.py:140-143)."""
    d = depth_frame.astype(np.int64)
    g = ((d[..., 0] * 3735 + d[..., 1] * 19235 + d[..., 2] * 9798 + (1 << 14)) >> 15).astype(np.uint8)
    return (torch.from_numpy(g).to(torch_device).float().unsqueeze(0)) / 255.0


def tensor_to_frame(tensor):
    """RGB f32 CHW -> BGR u8 HWC, truncating (core/render_3d.py:289-291)."""
    a = (tensor.permute(1, 2, 0).cpu().numpy() * 255).astype(np.uint8)
    return np.ascontiguousarray(a[..., ::-1])


def _ptr_and_mem(x):
    """(address, mem kind, keepalive) for a torch tensor or numpy array."""
    if torch is not None and isinstance(x, torch.Tensor):
        x = x.contiguous()
        if x.is_cuda:
            return x.data_ptr(), _lib.MEM_DEVICE, x
        a = x.numpy()
        return a.ctypes.data, _lib.MEM_HOST, a
    a = np.ascontiguousarray(x)
    return a.ctypes.data, _lib.MEM_HOST, a


def pixel_shift_cuda(
    frame_tensor, depth_tensor, width, height, fg_shift, mg_shift, bg_shift,
    blur_ksize=9, feather_strength=10.0, max_pixel_shift_percent=0.02, parallax_balance=0.8,
    zero_parallax_strength=0.0, use_subject_tracking=True, enable_floating_window=True,
    return_shift_map=True, enable_feathering=True, enable_edge_masking=True, dof_strength=2.0,
    convergence_strength=0.0, enable_dynamic_convergence=True, depth_pop_gamma=0.85,
    depth_pop_mid=0.50, depth_stretch_lo=0.05, depth_stretch_hi=0.95, fg_pop_multiplier=1.20,
    bg_push_multiplier=1.10, subject_lock_strength=1.00, _info=None,
):
    """core/render_3d.py:561-712.  frame_tensor f32 RGB [3,h,w] in 0..1, depth_tensor f32
    [1,h,w]; returns (left_bgr_u8, right_bgr_u8[, final_shift cpu f32 [1,H,W]]).
    `dof_strength` is accepted and unused, as in the reference."""
    width, height = int(width), int(height)
    ctx = _ctx()
    fp, fmem, fk = _ptr_and_mem(frame_tensor if frame_tensor.dtype in (np.float32, getattr(torch, "float32", None))
                                else frame_tensor.float())
    dp, dmem, dk = _ptr_and_mem(depth_tensor if depth_tensor.dtype in (np.float32, getattr(torch, "float32", None))
                                else depth_tensor.float())
    in_h, in_w = int(frame_tensor.shape[1]), int(frame_tensor.shape[2])
    assert tuple(depth_tensor.shape[-2:]) == (in_h, in_w), "Shape mismatch"
    if fmem != dmem:
        raise ValueError("frame and depth must live on the same device")
    p = ShiftParams(
        float(fg_shift), float(mg_shift), float(bg_shift), int(blur_ksize), float(feather_strength),
        float(max_pixel_shift_percent), float(parallax_balance), float(zero_parallax_strength),
        int(bool(use_subject_tracking)), int(bool(enable_floating_window)), int(bool(enable_feathering)),
        int(bool(enable_edge_masking)), float(convergence_strength), int(bool(enable_dynamic_convergence)),
        float(depth_pop_gamma), float(depth_pop_mid), float(depth_stretch_lo), float(depth_stretch_hi),
        float(fg_pop_multiplier), float(bg_push_multiplier), float(subject_lock_strength))
    info = FrameInfo()
    if fmem == _lib.MEM_DEVICE:
        left_t = torch.empty((height, width, 3), dtype=torch.uint8, device=frame_tensor.device)
        right_t = torch.empty_like(left_t)
        shift_t = torch.empty((1, height, width), dtype=torch.float32, device=frame_tensor.device) \
            if return_shift_map else None
        torch.cuda.current_stream().synchronize()
        ctx.check(ctx.lib.vd3d_pixel_shift(
            ctx.h, fp, dp, in_h, in_w, width, height, C.byref(p), left_t.data_ptr(), right_t.data_ptr(),
            shift_t.data_ptr() if shift_t is not None else None, _lib.MEM_DEVICE, C.byref(info)))
        left, right = left_t.cpu().numpy(), right_t.cpu().numpy()
        shift = shift_t.cpu() if shift_t is not None else None
    else:
        left = np.empty((height, width, 3), dtype=np.uint8)
        right = np.empty_like(left)
        sh = np.empty((1, height, width), dtype=np.float32) if return_shift_map else None
        ctx.check(ctx.lib.vd3d_pixel_shift(
            ctx.h, fp, dp, in_h, in_w, width, height, C.byref(p), left.ctypes.data, right.ctypes.data,
            sh.ctypes.data if sh is not None else None, _lib.MEM_HOST, C.byref(info)))
        shift = (torch.from_numpy(sh) if torch is not None else sh) if sh is not None else None
    if _info is not None:
        _info.append(info)
    if return_shift_map:
        return left, right, shift
    return left, right


# ---------------------------------------------------------------------------
# stage functions (same kernels as the fused frame path)
# ---------------------------------------------------------------------------
def apply_sharpening(frame, factor=1.0):
    """core/render_3d.py:717-732 (cv2.filter2D u8 semantics)."""
    ctx = _ctx()
    src = np.ascontiguousarray(frame, dtype=np.uint8)
    out = np.empty_like(src)
    ctx.check(ctx.lib.vd3d_sharpen(ctx.h, src.ctypes.data, src.shape[0], src.shape[1], float(factor),
                                   out.ctypes.data, _lib.MEM_HOST))
    return out


def dof_grade_frame(frame_bgr, depth01, focal_depth, max_sigma=2.0, saturation=1.0, contrast=1.0,
                    brightness=0.0):
    """frame_to_tensor -> apply_dof_cuda -> apply_color_grade -> tensor_to_frame on a u8 eye
    (core/render_3d.py:1342-1369); max_sigma <= 0 skips DOF (1373-1386)."""
    ctx = _ctx()
    src = np.ascontiguousarray(frame_bgr, dtype=np.uint8)
    out = np.empty_like(src)
    d = None
    dh = dw = 0
    if depth01 is not None:
        d = np.ascontiguousarray(np.asarray(depth01, dtype=np.float32).reshape(depth01.shape[-2:]))
        dh, dw = d.shape
    ctx.check(ctx.lib.vd3d_dof_grade(
        ctx.h, src.ctypes.data, src.shape[0], src.shape[1], d.ctypes.data if d is not None else None, dh, dw,
        float(focal_depth), float(max_sigma), float(saturation), float(contrast), float(brightness),
        out.ctypes.data, _lib.MEM_HOST))
    return out


def heal_missing_pixels(warped_frame, warped_depth, original_frame, edge_mask, heal_strength=0.5):
    """core/render_3d.py:431-459 (`warped_depth` is unused there too).  f32 [3,H,W] tensors / arrays in,
    same type out."""
    ctx = _ctx()
    is_t = torch is not None and isinstance(warped_frame, torch.Tensor)
    wf = np.ascontiguousarray((warped_frame.detach().cpu().numpy() if is_t else warped_frame), dtype=np.float32)
    of = original_frame.detach().cpu().numpy() if (torch is not None and isinstance(original_frame, torch.Tensor)) \
        else original_frame
    of = np.ascontiguousarray(of, dtype=np.float32)
    em = None
    if edge_mask is not None:
        em = edge_mask.detach().cpu().numpy() if (torch is not None and isinstance(edge_mask, torch.Tensor)) else edge_mask
        em = np.ascontiguousarray(np.asarray(em, dtype=np.float32).reshape(wf.shape[1:]))
    out = np.empty_like(wf)
    ctx.check(ctx.lib.vd3d_heal(ctx.h, wf.ctypes.data, of.ctypes.data, em.ctypes.data if em is not None else None,
                                wf.shape[1], wf.shape[2], float(heal_strength), out.ctypes.data, _lib.MEM_HOST))
    return torch.from_numpy(out).to(warped_frame.device) if is_t else out


def _fit_eye(image, target_width, target_height, keep_aspect):
    ctx = _ctx()
    img = np.ascontiguousarray(image, dtype=np.uint8)
    assert img.ndim == 3 and img.shape[2] == 3, "expected a BGR u8 image"
    h, w = img.shape[:2]
    out = np.empty((int(target_height), int(target_width), 3), dtype=np.uint8)
    ctx.check(ctx.lib.vd3d_fit_eye(ctx.h, img.ctypes.data, h, w, int(target_width), int(target_height),
                                   int(bool(keep_aspect)), out.ctypes.data, _lib.MEM_HOST))
    return out


def pad_to_aspect_ratio(image, target_width, target_height, bg_color=(0, 0, 0)):
    """core/render_3d.py:101-131 on the GPU (vd3d_fit_eye): aspect-preserving cv2 INTER_AREA shrink (identity, integer
    or fractional factors) centred on a black canvas.  Enlarging fits raise (cv2 changes algorithm there)."""
    if tuple(int(c) for c in bg_color) != (0, 0, 0):
        raise NotImplementedError("pad_to_aspect_ratio: only the black canvas the render loop uses is supported")
    return _fit_eye(image, target_width, target_height, True)


def resize_area(image, width, height):
    """cv2.resize(image, (width, height), interpolation=cv2.INTER_AREA) as in the Half-SBS eye fit
    (core/render_3d.py:1413-1414), shrinking only."""
    return _fit_eye(image, width, height, False)


def _pack(left, right, fmt):
    """format_3d_output on the GPU (vd3d_pack): SBS hstack, Dubois anaglyph, row interlace."""
    ctx = _ctx()
    l = np.ascontiguousarray(left, dtype=np.uint8)
    r = np.ascontiguousarray(right, dtype=np.uint8)
    assert l.shape == r.shape and l.ndim == 3 and l.shape[2] == 3, "Shape mismatch"
    h, w = l.shape[:2]
    code = _lib.FMT[fmt]
    out = np.empty((h, 2 * w, 3) if code in (0, 1, 4) else (h, w, 3), dtype=np.uint8)
    ctx.check(ctx.lib.vd3d_pack(ctx.h, l.ctypes.data, r.ctypes.data, h, w, code, out.ctypes.data, _lib.MEM_HOST))
    return out


def generate_anaglyph_3d(left_frame, right_frame):
    """core/render_3d.py:862-883 (Dubois matrix on channel indices 0,1,2 as written there)."""
    return _pack(left_frame, right_frame, "Red-Cyan Anaglyph")


def format_3d_output(left, right, fmt):
    """core/render_3d.py:837-860."""
    if fmt == "VR" and tuple(left.shape[:2]) != (1600, 1440):
        # the render loop always hands 1440x1600 eyes (pad_to_aspect_ratio) where the reference's cv2.resize is the identity
        raise NotImplementedError("VR pack of eyes that are not 1440x1600 (cv2 INTER_LINEAR resize) is outside the hot path")
    if fmt not in ("Half-SBS", "Full-SBS", "Red-Cyan Anaglyph", "Passive Interlaced", "VR"):
        fmt = "Half-SBS"  # the reference falls back to hstack (860)
    return _pack(left, right, fmt)


# ---------------------------------------------------------------------------
# frame loop
# ---------------------------------------------------------------------------
def make_render_params(output_width, output_height, fg_shift, mg_shift, bg_shift, sharpness_factor,
                       output_format, aspect_ratio, dof_strength, feather_strength=0.0, blur_ksize=1,
                       use_subject_tracking=False, use_floating_window=False, max_pixel_shift_percent=0.02,
                       preserve_original_aspect=False, zero_parallax_strength=0.0, enable_edge_masking=True,
                       enable_feathering=True, original_video_width=None, original_video_height=None,
                       convergence_strength=0.0, enable_dynamic_convergence=True, ipd_factor=1.0,
                       color_saturation=1.0, color_contrast=1.0, color_brightness=0.0):
    """vd3d_render_params from render_sbs_3d's arguments (core/render_3d.py:933-985)."""
    if output_format not in _lib.FMT:
        output_format = "Half-SBS"  # format_3d_output's fallback is hstack (860)
    return RenderParams(
        int(output_width), int(output_height), float(fg_shift), float(mg_shift), float(bg_shift),
        float(sharpness_factor), _lib.FMT[output_format], float(aspect_ratio), float(dof_strength),
        float(feather_strength), int(blur_ksize), int(bool(use_subject_tracking)),
        int(bool(use_floating_window)), float(max_pixel_shift_percent), int(bool(preserve_original_aspect)),
        float(zero_parallax_strength), int(bool(enable_edge_masking)), int(bool(enable_feathering)),
        int(original_video_width or 0), int(original_video_height or 0), float(convergence_strength),
        int(bool(enable_dynamic_convergence)), float(ipd_factor), float(color_saturation),
        float(color_contrast), float(color_brightness))


def plan_sizes(src_w, src_h, rp):
    pl = SizePlan()
    rc = _lib.load().vd3d_plan_sizes(int(src_w), int(src_h), C.byref(rp), C.byref(pl))
    if rc != 0:
        raise _lib.Vd3dError(f"vd3d_plan_sizes: unsupported configuration ({rc})")
    return pl


def output_shape(rp, pl):
    if rp.output_format in (_lib.FMT["Red-Cyan Anaglyph"], _lib.FMT["Passive Interlaced"]):
        return (pl.per_eye_h, pl.per_eye_w, 3)
    return (pl.out_height, pl.out_width, 3)


def render_frame(frame_bgr, depth_bgr, rp, want_info=False, ctx=None):
    """One iteration of the render_sbs_3d loop body (core/render_3d.py:1227-1419) on host
    uint8 frames; returns the packed BGR frame (and vd3d_frame_info)."""
    ctx = ctx or _ctx()
    f = np.ascontiguousarray(frame_bgr, dtype=np.uint8)
    d = np.ascontiguousarray(depth_bgr, dtype=np.uint8)
    dch = 1 if d.ndim == 2 else d.shape[2]
    sh, sw = f.shape[:2]
    pl = plan_sizes(sw, sh, rp)
    out = np.empty(output_shape(rp, pl), dtype=np.uint8)
    info = FrameInfo()
    ctx.check(ctx.lib.vd3d_render_frame(ctx.h, f.ctypes.data, d.ctypes.data, dch, sh, sw, C.byref(rp),
                                        out.ctypes.data, _lib.MEM_HOST, C.byref(info) if want_info else None))
    return (out, info) if want_info else out


def advance_state(frame_bgr, depth_bgr, rp, ctx=None):
    """One loop iteration without rendering (temporal state only): the building block of exact
    multi-GPU frame sharding (visiondepth3d_b200/sharding.py)."""
    ctx = ctx or _ctx()
    f = np.ascontiguousarray(frame_bgr, dtype=np.uint8)
    d = np.ascontiguousarray(depth_bgr, dtype=np.uint8)
    dch = 1 if d.ndim == 2 else d.shape[2]
    ctx.check(ctx.lib.vd3d_advance_state(ctx.h, f.ctypes.data, d.ctypes.data, dch, f.shape[0], f.shape[1],
                                         C.byref(rp), _lib.MEM_HOST))


def render_sbs_3d(
    input_path, depth_path, output_path, selected_codec, fps, output_width, output_height,
    fg_shift, mg_shift, bg_shift, sharpness_factor, output_format, selected_aspect_ratio,
    aspect_ratios, dof_strength, feather_strength=0.0, blur_ksize=1, use_ffmpeg=False,
    selected_ffmpeg_codec=None, crf_value=23, use_subject_tracking=False, use_floating_window=False,
    max_pixel_shift_percent=0.02, progress=None, progress_label=None, suspend_flag=None,
    cancel_flag=None, auto_crop_black_bars=False, parallax_balance=0.8,
    preserve_original_aspect=False, zero_parallax_strength=0.0, enable_edge_masking=True,
    enable_feathering=True, skip_blank_frames=False, original_video_width=None,
    original_video_height=None, convergence_strength=0.0, enable_dynamic_convergence=True,
    ipd_factor=1.0, depth_pop_gamma=0.85, depth_pop_mid=0.50, depth_stretch_lo=0.05,
    depth_stretch_hi=0.95, fg_pop_multiplier=1.20, bg_push_multiplier=1.10,
    subject_lock_strength=1.00, color_saturation=1.0, color_contrast=1.0, color_brightness=0.0,
    start_s=None, end_s=None,
):
    """core/render_3d.py:933-1504: video in -> video out.  Video I/O stays cv2 on the host
    (SURVEY section 8(f) "next"); every frame's math runs in libvd3d.  Reproduces the reference's
    sequencing: first frame of the clip window dropped (1184-1188), pop controls and
    parallax_balance not forwarded (1284-1331), no exception escapes (1477-1478)."""
    import cv2
    if use_ffmpeg or auto_crop_black_bars or skip_blank_frames:
        print("⚠️ use_ffmpeg / auto_crop_black_bars / skip_blank_frames are outside the B200 hot path")
        return
    cap, dcap = cv2.VideoCapture(input_path), cv2.VideoCapture(depth_path)
    if not cap.isOpened() or not dcap.isOpened():
        return
    total_frames_full = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    fps = cap.get(cv2.CAP_PROP_FPS) or fps or 30.0
    dur_ms = (total_frames_full / max(fps, 1e-6)) * 1000.0
    start_ms = max(0.0, (start_s or 0.0) * 1000.0)
    end_ms = dur_ms if (end_s is None) else min(dur_ms, end_s * 1000.0)
    if start_ms >= end_ms - 0.5:
        print("⚠️ Invalid clip window; nothing to render.")
        cap.release(); dcap.release()
        return
    start_frame_idx = int(round(start_ms / 1000.0 * fps))
    end_frame_idx = int(round(end_ms / 1000.0 * fps))
    clip_total_frames = max(0, end_frame_idx - start_frame_idx)
    cap.set(cv2.CAP_PROP_POS_FRAMES, start_frame_idx)
    dcap.set(cv2.CAP_PROP_POS_FRAMES, start_frame_idx)
    ret1, frame = cap.read()
    ret2, depth = dcap.read()
    if not ret1 or not ret2:
        cap.release(); dcap.release()
        return
    suspend_flag = suspend_flag or globals()["suspend_flag"]
    cancel_flag = cancel_flag or globals()["cancel_flag"]
    ratio = aspect_ratios.get(selected_aspect_ratio.get(), 16 / 9)
    rp = make_render_params(
        output_width, output_height, fg_shift, mg_shift, bg_shift, sharpness_factor, output_format, ratio,
        dof_strength, feather_strength, blur_ksize, use_subject_tracking, use_floating_window,
        max_pixel_shift_percent, preserve_original_aspect, zero_parallax_strength, enable_edge_masking,
        enable_feathering, original_video_width, original_video_height, convergence_strength,
        enable_dynamic_convergence, ipd_factor, color_saturation, color_contrast, color_brightness)
    sh, sw = frame.shape[:2]
    out = None
    try:
        pl = plan_sizes(sw, sh, rp)
        oshape = output_shape(rp, pl)
        out = cv2.VideoWriter(output_path, cv2.VideoWriter_fourcc(*selected_codec), fps, (oshape[1], oshape[0]))
        if not out.isOpened():
            print("❌ OpenCV VideoWriter failed to open. Check codec/fourcc and path.")
            cap.release(); dcap.release()
            return
        ctx = _ctx()
        ctx.reset(_lib.STATE_CLIP)  # ShiftSmoother / TemporalDepthFilter / FocalDepthTracker are per render
        cap.set(cv2.CAP_PROP_POS_FRAMES, start_frame_idx)
        dcap.set(cv2.CAP_PROP_POS_FRAMES, start_frame_idx)
        ret1, frame = cap.read()   # read + discard (1184-1188)
        ret2, depth = dcap.read()
        if not ret1 or not ret2:
            return
        total_frames = clip_total_frames if clip_total_frames > 0 else total_frames_full
        t0 = time.time()
        for idx in range(total_frames):
            if cancel_flag.is_set():
                break
            while suspend_flag.is_set() and not cancel_flag.is_set():
                time.sleep(0.2)
            if cancel_flag.is_set():
                break
            ret1, frame = cap.read()
            ret2, depth = dcap.read()
            if not ret1 or not ret2:
                break
            final = render_frame(frame, depth, rp, ctx=ctx)
            out.write(final)
            if end_s is not None and int(cap.get(cv2.CAP_PROP_POS_FRAMES)) >= end_frame_idx:
                break
            if progress:
                progress["value"] = (idx / max(total_frames, 1)) * 100.0
                progress.update()
            if progress_label:
                el = time.time() - t0
                progress_label.config(text=f"{(idx / max(total_frames, 1)) * 100.0:.2f}% | "
                                           f"FPS: {(idx + 1) / max(el, 1e-9):.2f}")
        if progress:
            progress["value"] = 100
            progress.update()
    except Exception as e:  # the reference prints and returns None (1477-1478)
        print(f"❌ Render crashed: {e}")
    finally:
        cap.release(); dcap.release()
        if out is not None:
            try:
                out.release()
            except Exception:
                pass
