"""Prepare Depth-Anything-V2 weights (HF state_dict naming, transformers 5.5
DepthAnythingForDepthEstimation) for the libvd3d depth engine: f16 K-major GEMM operands,
3x3 conv kernels as [Cout, tap, Cin_pad], ConvTranspose as pixel-shuffle GEMMs, position
embeddings interpolated for the processed size.  Pure set-up code (runs once per model)."""
import numpy as np
import torch
import torch.nn.functional as F

CONFIGS = {  # SURVEY section 8(c); verify against config.json when real checkpoints are supplied
    "vits": dict(hidden=384, layers=12, heads=6, taps=[3, 6, 9, 12], neck=[48, 96, 192, 384], fusion=64),
    "vitb": dict(hidden=768, layers=12, heads=12, taps=[3, 6, 9, 12], neck=[96, 192, 384, 768], fusion=128),
    "vitl": dict(hidden=1024, layers=24, heads=16, taps=[5, 12, 18, 24], neck=[256, 512, 1024, 1024], fusion=256),
}


def hf_config(name):
    """transformers config objects equivalent to depth-anything/Depth-Anything-V2-{Small,Base,Large}-hf."""
    from transformers import DepthAnythingConfig, Dinov2Config
    c = CONFIGS[name]
    bc = Dinov2Config(hidden_size=c["hidden"], num_hidden_layers=c["layers"], num_attention_heads=c["heads"],
                      image_size=518, patch_size=14, out_indices=c["taps"], apply_layernorm=True,
                      reshape_hidden_states=False)
    return DepthAnythingConfig(backbone_config=bc, reassemble_hidden_size=c["hidden"],
                               neck_hidden_sizes=c["neck"], fusion_hidden_size=c["fusion"], head_hidden_size=32,
                               patch_size=14, reassemble_factors=[4, 2, 1, 0.5])


def _up(v, m):
    return (v + m - 1) // m * m


def _f16(t):
    return np.ascontiguousarray(t.detach().to(torch.float32).cpu().numpy().astype(np.float16))


def _f32(t):
    return np.ascontiguousarray(t.detach().to(torch.float32).cpu().numpy())


def _conv3(w, cin_pad, cout_pad=None):
    """[Cout, Cin, 3, 3] -> [Cout_pad, 9 * cin_pad] (tap-major, channel-minor), zero padded."""
    co, ci = w.shape[0], w.shape[1]
    cout_pad = cout_pad or co
    out = torch.zeros(cout_pad, 9, cin_pad, dtype=torch.float32)
    out[:co, :, :ci] = w.detach().float().permute(0, 2, 3, 1).reshape(co, 9, ci)
    return _f16(out.reshape(cout_pad, 9 * cin_pad))


def _pad_vec(b, n):
    out = torch.zeros(n, dtype=torch.float32)
    out[: b.numel()] = b.detach().float().reshape(-1)
    return _f32(out)


def prepare(sd, cfg, image_h, image_w):
    """state_dict -> {name: np.ndarray} in the layouts depth_engine.cu expects."""
    D, L, Fz = cfg["hidden"], cfg["layers"], cfg["fusion"]
    ph, pw = image_h // 14, image_w // 14
    out = {}
    e = "backbone.embeddings."
    w = torch.zeros(D, 592)
    w[:, :588] = sd[e + "patch_embeddings.projection.weight"].float().reshape(D, 588)
    out["pe.w"] = _f16(w)
    out["pe.b"] = _f32(sd[e + "patch_embeddings.projection.bias"])
    out["cls"] = _f32(sd[e + "cls_token"].reshape(D))
    pos = sd[e + "position_embeddings"].float()
    npos = pos.shape[1] - 1
    g = int(round(npos ** 0.5))
    if not (ph == g and pw == g):
        pp = pos[:, 1:].reshape(1, g, g, D).permute(0, 3, 1, 2)
        pp = F.interpolate(pp, size=(ph, pw), mode="bicubic", align_corners=False)  # Dinov2Embeddings.interpolate_pos_encoding
        pos = torch.cat((pos[:, :1], pp.permute(0, 2, 3, 1).reshape(1, -1, D)), dim=1)
    out["pos"] = _f32(pos.reshape(-1, D))
    for i in range(L):
        p = f"backbone.encoder.layer.{i}."
        a = p + "attention.attention."
        out[f"l{i}.ln1.g"] = _f32(sd[p + "norm1.weight"])
        out[f"l{i}.ln1.b"] = _f32(sd[p + "norm1.bias"])
        out[f"l{i}.qkv.w"] = _f16(torch.cat([sd[a + "query.weight"], sd[a + "key.weight"], sd[a + "value.weight"]], 0))
        out[f"l{i}.qkv.b"] = _f32(torch.cat([sd[a + "query.bias"], sd[a + "key.bias"], sd[a + "value.bias"]], 0))
        out[f"l{i}.proj.w"] = _f16(sd[p + "attention.output.dense.weight"])
        out[f"l{i}.proj.b"] = _f32(sd[p + "attention.output.dense.bias"])
        out[f"l{i}.ls1"] = _f32(sd[p + "layer_scale1.lambda1"])
        out[f"l{i}.ln2.g"] = _f32(sd[p + "norm2.weight"])
        out[f"l{i}.ln2.b"] = _f32(sd[p + "norm2.bias"])
        out[f"l{i}.fc1.w"] = _f16(sd[p + "mlp.fc1.weight"])
        out[f"l{i}.fc1.b"] = _f32(sd[p + "mlp.fc1.bias"])
        out[f"l{i}.fc2.w"] = _f16(sd[p + "mlp.fc2.weight"])
        out[f"l{i}.fc2.b"] = _f32(sd[p + "mlp.fc2.bias"])
        out[f"l{i}.ls2"] = _f32(sd[p + "layer_scale2.lambda1"])
    out["norm.g"] = _f32(sd["backbone.layernorm.weight"])
    out["norm.b"] = _f32(sd["backbone.layernorm.bias"])
    for i, C in enumerate(cfg["neck"]):
        CP = _up(C, 64)
        r = f"neck.reassemble_stage.layers.{i}."
        pw_ = torch.zeros(CP, D)
        pw_[:C] = sd[r + "projection.weight"].float().reshape(C, D)
        out[f"r{i}.proj.w"] = _f16(pw_)
        out[f"r{i}.proj.b"] = _pad_vec(sd[r + "projection.bias"], CP)
        if i < 2:  # ConvTranspose2d [Cin, Cout, k, k] -> rows n = (dy*k+dx)*CP + co, cols ci
            k = 4 if i == 0 else 2
            wt = sd[r + "resize.weight"].float()
            uw = torch.zeros(k * k, CP, CP)
            uw[:, :C, :C] = wt.permute(2, 3, 1, 0).reshape(k * k, C, C)
            out[f"r{i}.up.w"] = _f16(uw.reshape(k * k * CP, CP))
            ub = torch.zeros(k * k, CP)
            ub[:, :C] = sd[r + "resize.bias"].float()[None, :]
            out[f"r{i}.up.b"] = _f32(ub.reshape(-1))
        elif i == 3:
            out["r3.down.w"] = _conv3(sd[r + "resize.weight"], CP, CP)
            out["r3.down.b"] = _pad_vec(sd[r + "resize.bias"], CP)
        out[f"n{i}.conv.w"] = _conv3(sd[f"neck.convs.{i}.weight"], CP)
    for j in range(4):
        f = f"neck.fusion_stage.layers.{j}."
        for unit, hf in (("rl1", "residual_layer1"), ("rl2", "residual_layer2")):
            for c, hc in (("c1", "convolution1"), ("c2", "convolution2")):
                out[f"f{j}.{unit}.{c}.w"] = _conv3(sd[f + hf + "." + hc + ".weight"], Fz)
                out[f"f{j}.{unit}.{c}.b"] = _f32(sd[f + hf + "." + hc + ".bias"])
        out[f"f{j}.proj.w"] = _f16(sd[f + "projection.weight"].reshape(Fz, Fz))
        out[f"f{j}.proj.b"] = _f32(sd[f + "projection.bias"])
    F2 = _up(Fz // 2, 64)
    out["h.c1.w"] = _conv3(sd["head.conv1.weight"], Fz, F2)
    out["h.c1.b"] = _pad_vec(sd["head.conv1.bias"], F2)
    out["h.c2.w"] = _conv3(sd["head.conv2.weight"], F2)
    out["h.c2.b"] = _f32(sd["head.conv2.bias"])
    out["h.c3.w"] = _f32(sd["head.conv3.weight"].reshape(-1))
    out["h.c3.b"] = _f32(sd["head.conv3.bias"].reshape(-1))
    return out
