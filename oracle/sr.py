"""oracle/sr.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the Real-ESRGAN upscale stage of the reference's FrameTools pipeline
(core/merged_pipeline.py:221-284: preprocess_esr -> ONNX session -> postprocess_esr -> the INTER_CUBIC resize chain ->
blend_images).

The network itself is NOT in /root/reference: the reference runs an ONNX file it does not ship
(`weights/RealESR_Gx4_fp16.onnx`, VisionDepth3D.py:1094-1100).  That file is the export of
realesr-general-x4v3 from xinntao/Real-ESRGAN v0.3.0 (`realesrgan/archs/srvgg_arch.py`, SRVGGNetCompact with
num_feat=64, num_conv=32, upscale=4, PReLU; `RealESR_Animex4` = realesr-animevideov3, the same class with
num_conv=16).  Its published architecture is restated here on the upstream state_dict naming (`body.{i}.weight`):
    conv3x3(3->64) PReLU, num_conv x [conv3x3(64->64) PReLU], conv3x3(64->48), PixelShuffle(4), + nearest x4 of the input.
PARITY UNPINNED for the network: neither the upstream package nor the ONNX file exists offline, so the restatement is
checked only against itself (shape / pixel-shuffle / base-add identities) and used with random-init weights.  The
pre/post-processing chain around it IS pinned against cv2 (tests/test_oracle_sr.py).
"""
import numpy as np


def srvgg_state_dict(num_conv=32, num_feat=64, seed=0):
    """Random-init SRVGGNetCompact weights in the upstream naming (no checkpoints offline).  Convolution weights are
    scaled so that activations stay O(1) through the stack, like a trained model's."""
    import torch
    g = torch.Generator().manual_seed(seed)
    sd = {}
    chans = [(3, num_feat)] + [(num_feat, num_feat)] * num_conv + [(num_feat, 48)]
    for i, (ci, co) in enumerate(chans):
        std = (2.0 / (1.0 + 0.25 ** 2) / (9 * ci)) ** 0.5          # He init for PReLU(0.25)
        if i == len(chans) - 1:
            std *= 0.1                                            # the residual on top of the nearest-upsampled input
        sd[f"body.{2 * i}.weight"] = torch.randn(co, ci, 3, 3, generator=g) * std
        sd[f"body.{2 * i}.bias"] = torch.randn(co, generator=g) * 0.02
        if i < len(chans) - 1:
            sd[f"body.{2 * i + 1}.weight"] = 0.25 + 0.1 * torch.rand(co, generator=g)   # PReLU slopes
    return sd


def srvgg_forward(sd, x):
    """x f32 [1,3,h,w] in 0..1 (RGB) -> f32 [1,3,4h,4w] (SRVGGNetCompact.forward, srvgg_arch.py)."""
    import torch
    import torch.nn.functional as F
    n_conv = len([k for k in sd if k.endswith(".bias")])
    out = x
    for i in range(n_conv):
        out = F.conv2d(out, sd[f"body.{2 * i}.weight"], sd[f"body.{2 * i}.bias"], padding=1)
        if i < n_conv - 1:
            out = F.prelu(out, sd[f"body.{2 * i + 1}.weight"])
    out = F.pixel_shuffle(out, 4)
    return out + F.interpolate(x, scale_factor=4, mode="nearest")


def preprocess_esr(frame_bgr):
    """core/merged_pipeline.py:221-225."""
    img = frame_bgr[..., ::-1].astype(np.float32) / np.float32(255.0)
    return np.ascontiguousarray(np.transpose(img, (2, 0, 1))[None]).astype(np.float32)


def postprocess_esr(tensor):
    """core/merged_pipeline.py:227-231: clip to [0,1], x255, truncate, RGB -> BGR."""
    t = np.transpose(np.squeeze(tensor, axis=0), (1, 2, 0))
    t = np.clip(t, 0, 1) * np.float32(255.0)
    return np.ascontiguousarray(t.astype(np.uint8)[..., ::-1])


def resize_cubic_bgr(img, ow, oh):
    """cv2.resize(BGR u8, (ow, oh), interpolation=cv2.INTER_CUBIC): per channel the float bicubic of
    oracle.dibr.resize_cubic_u8 (the installed cv2 / IPP arithmetic)."""
    from oracle.dibr import resize_cubic_u8
    if img.shape[:2] == (oh, ow):
        return img.copy()
    return np.stack([resize_cubic_u8(np.ascontiguousarray(img[..., c]), ow, oh) for c in range(img.shape[2])], axis=-1)


def blend_images(original, upscaled, mode="OFF"):
    """core/merged_pipeline.py:233-238: cv2.addWeighted(upscaled, a, original, 1-a, 0).  cv2 4.13 evaluates it per
    byte in float32 as fma(upscaled, a, fl(original * (1 - a))), then rounds half to even and saturates (found by
    search against the real op, tests/test_oracle_sr.py: exact)."""
    if mode == "OFF":
        return upscaled
    alpha = {"LOW": 0.85, "MEDIUM": 0.5, "HIGH": 0.25}.get(mode.upper(), 1.0)
    a, b = np.float32(alpha), np.float32(1 - alpha)
    t = (original.astype(np.float32) * b).astype(np.float64)                       # fl32(o * beta)
    v = (upscaled.astype(np.float64) * float(a) + t).astype(np.float32)            # one rounding: the fma
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def run_esrgan(sd, frame_bgr, blend_mode="OFF", input_res_pct=100, target_size=None):
    """core/merged_pipeline.py:240-267 without tiling: optional input resize (INTER_AREA below 100 %, INTER_CUBIC
    above), network, postprocess, resize to 4x (identity), back to the ORIGINAL size, optionally to target_size, blend."""
    import torch
    from oracle.dibr import resize_area
    original = frame_bgr
    frame = frame_bgr
    if input_res_pct != 100:
        h, w = frame.shape[:2]
        nw, nh = int(w * input_res_pct / 100), int(h * input_res_pct / 100)
        frame = resize_area(frame, nw, nh) if input_res_pct < 100 else resize_cubic_bgr(frame, nw, nh)
    with torch.no_grad():
        out = srvgg_forward(sd, torch.from_numpy(preprocess_esr(frame))).numpy()
    up = postprocess_esr(out)
    up = resize_cubic_bgr(up, frame.shape[1] * 4, frame.shape[0] * 4)
    up = resize_cubic_bgr(up, original.shape[1], original.shape[0])
    if target_size:
        up = resize_cubic_bgr(up, target_size[0], target_size[1])
    return blend_images(original, up, blend_mode)
