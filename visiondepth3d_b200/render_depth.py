"""Drop-in for the hot-path surface of the reference's core/render_depth.py: the `pipe`
callable protocol and the depth -> uint8 normalisation, backed by the libvd3d depth engine
(Depth-Anything-V2 on tcgen05 tensor cores).  GUI, ONNX, Marigold, DepthCrafter and the
letterbox tracker are outside the B200 hot path.

Protocol kept (core/render_depth.py:1113-1119, consumers 201-268 and 1894-1917):
    pipe(images: list[PIL.Image], inference_size: (W, H) | None) -> list[{"predicted_depth": Tensor[h, w]}]
with `predicted_depth` already resized (bicubic) to the PIL image size, as
transformers' DepthEstimationPipeline.postprocess does.
"""
import numpy as np

from .depth_engine import DepthEngine
from .depth_weights import CONFIGS, hf_config

try:
    import torch
    torch.set_grad_enabled(False)  # the reference does this at import (core/render_depth.py:42)
except Exception:  # pragma: no cover
    torch = None

# module globals the reference's callers read (core/render_depth.py:34-36, 992)
pipe = None
pipe_type = None
_engine = None

# the three checkpoints of the hot path (core/render_depth.py:695-697)
supported_models = {
    "Depth Anything V2 Small": ("depth-anything/Depth-Anything-V2-Small-hf", "vits"),
    "Depth Anything V2 Base": ("depth-anything/Depth-Anything-V2-Base-hf", "vitb"),
    "Depth Anything V2 Large": ("depth-anything/Depth-Anything-V2-Large-hf", "vitl"),
}


def _processed_size(width, height, target=518, multiple=14):
    """DPTImageProcessor: keep_aspect_ratio, ensure_multiple_of=14 (image_processing_dpt.py)."""
    sh, sw = target / height, target / width
    if abs(1 - sw) < abs(1 - sh):
        sh = sw
    else:
        sw = sh
    rnd = lambda v: max(multiple, int(round(v / multiple) * multiple))  # noqa: E731
    return rnd(sh * height), rnd(sw * width)


def load_checkpoint(path):
    """HF-format checkpoint (model.safetensors / pytorch_model.bin of
    depth-anything/Depth-Anything-V2-{Small,Base,Large}-hf) -> state_dict.  If a config.json sits next
    to it, the architecture is cross-checked against depth_weights.CONFIGS."""
    import json
    import os
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        sd = load_file(path)
    else:
        sd = torch.load(path, map_location="cpu")
    arch = None
    hidden = sd["backbone.embeddings.cls_token"].shape[-1]
    for name, c in CONFIGS.items():
        if c["hidden"] == hidden:
            arch = name
    cfg_path = os.path.join(os.path.dirname(path), "config.json")
    if arch and os.path.exists(cfg_path):
        cj = json.load(open(cfg_path))
        c = CONFIGS[arch]
        got = (cj.get("neck_hidden_sizes"), cj.get("fusion_hidden_size"),
               cj.get("backbone_config", {}).get("out_indices"))
        want = (c["neck"], c["fusion"], c["taps"])
        if any(g is not None and list(g) != list(w) if isinstance(w, list) else (g is not None and g != w)
               for g, w in zip(got, want)):
            raise ValueError(f"config.json {got} does not match the built-in {arch} configuration {want}")
    return arch, sd


def load_depth_model(arch="vits", state_dict=None, width=1920, height=1080, seed=0):
    """Build the engine for frames of (width, height).  `state_dict` uses HF
    DepthAnythingForDepthEstimation naming (e.g. from a local safetensors checkpoint);
    without one a random-init model (torch.manual_seed(seed)) is used -- there is no network
    and the reference ships no weights."""
    global pipe, pipe_type, _engine
    if state_dict is None:
        from transformers import DepthAnythingForDepthEstimation
        torch.manual_seed(seed)
        state_dict = DepthAnythingForDepthEstimation(hf_config(arch)).eval().state_dict()
    ih, iw = _processed_size(width, height)
    _engine = DepthEngine(arch, ih, iw)
    _engine.load_state_dict(state_dict)
    pipe = hf_batch_safe_pipe
    pipe_type = "hf"
    return pipe, {"arch": arch, "processed_size": (ih, iw), "config": CONFIGS[arch]}


def hf_batch_safe_pipe(images, inference_size=None):
    """core/render_depth.py:1113-1119."""
    if _engine is None:
        raise RuntimeError("no depth model loaded: call load_depth_model() first")
    if not isinstance(images, (list, tuple)):
        images = [images]
    out = []
    for img in images:
        if inference_size is not None:
            from PIL import Image
            img = img.resize(tuple(inference_size), Image.BICUBIC)  # (1819-1821)
        rgb = np.asarray(img.convert("RGB"), dtype=np.uint8)
        want = _processed_size(rgb.shape[1], rgb.shape[0])
        if want != (_engine.image_h, _engine.image_w):
            raise ValueError(f"depth engine was built for processed size {(_engine.image_h, _engine.image_w)}, "
                             f"this image needs {want}: call load_depth_model(width=, height=) for this aspect")
        d32, _ = _engine.infer(np.ascontiguousarray(rgb[..., ::-1]))
        out.append({"predicted_depth": torch.from_numpy(d32) if torch is not None else d32})
    return out


def depth_u8_from_frame(frame_bgr, invert=False):
    """frame -> pipe -> convert_depth_to_grayscale in one GPU pass (u8 [h, w])."""
    if _engine is None:
        raise RuntimeError("no depth model loaded: call load_depth_model() first")
    return _engine.infer(frame_bgr, invert=invert)[1]


def convert_depth_to_grayscale(depth):
    """core/render_depth.py:585-611: PIL / tensor / ndarray, [H,W], [C,H,W] or [H,W,C] (C in {1,3}: first channel or the
    channel mean) -> per-frame min-max -> uint8 (truncating); NaN or flat frames give zeros, other types / ranks raise.
    Host helper for callers holding CPU data, like the reference's; the frame path does this on the GPU
    (k_depth_upsample_minmax + k_depth_to_u8)."""
    try:
        from PIL import Image
    except Exception:  # pragma: no cover
        Image = None
    if Image is not None and isinstance(depth, Image.Image):
        d = np.array(depth).astype(np.float32)
    elif torch is not None and isinstance(depth, torch.Tensor):
        d = depth.detach().cpu().float().numpy()
    elif isinstance(depth, np.ndarray):
        d = depth.astype(np.float32)
    else:
        raise TypeError(f"Unsupported depth type: {type(depth)}")
    if d.ndim == 3:
        if d.shape[0] in (1, 3):
            d = d[0] if d.shape[0] == 1 else d.mean(axis=0)
        elif d.shape[2] in (1, 3):
            d = d[..., 0] if d.shape[2] == 1 else d.mean(axis=-1)
    elif d.ndim != 2:
        raise ValueError(f"Unexpected depth shape: {d.shape}")
    lo, hi = np.min(d), np.max(d)
    if np.isnan(lo) or np.isnan(hi) or hi - lo < 1e-6:
        print("⚠️ Skipping frame with invalid depth values.")
        return np.zeros_like(d, dtype=np.uint8)
    return ((d - lo) / (hi - lo + 1e-6) * 255).astype(np.uint8)
