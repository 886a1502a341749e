"""oracle/depth.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

torch fp32 functional restatement of Depth-Anything-V2 as the reference runs it:
core/render_depth.py:1106-1119 builds transformers' pipeline("depth-estimation"), whose
model is DepthAnythingForDepthEstimation (third-party: transformers, unpinned in the
reference's requirements.txt:8; pinned here to 5.5.0 -- models/dinov2/modeling_dinov2.py,
models/depth_anything/modeling_depth_anything.py).  The algorithm lives outside
/root/reference, so it is restated from the published architecture directly on the HF
state_dict and pinned against the installed transformers module in
tests/test_oracle_depth.py (random-init weights: no checkpoints ship with the reference).
"""
import torch
import torch.nn.functional as F


def forward(sd, cfg, pixel_values, return_parts=False):
    """sd: HF state_dict; cfg: dict(hidden, layers, heads, taps, neck, fusion);
    pixel_values [3, H, W] f32 (H, W multiples of 14) -> predicted_depth [H, W]."""
    D, L, Hh = cfg["hidden"], cfg["layers"], cfg["heads"]
    x = pixel_values[None].float()
    _, _, IH, IW = x.shape
    ph, pw = IH // 14, IW // 14
    e = "backbone.embeddings."
    t = F.conv2d(x, sd[e + "patch_embeddings.projection.weight"], sd[e + "patch_embeddings.projection.bias"], stride=14)
    t = t.flatten(2).transpose(1, 2)  # [1, N, D]
    t = torch.cat((sd[e + "cls_token"], t), dim=1)
    pos = sd[e + "position_embeddings"]
    g = int(round((pos.shape[1] - 1) ** 0.5))
    if not (ph == g and pw == g):
        pp = pos[:, 1:].reshape(1, g, g, D).permute(0, 3, 1, 2)
        pp = F.interpolate(pp, size=(ph, pw), mode="bicubic", align_corners=False)
        pos = torch.cat((pos[:, :1], pp.permute(0, 2, 3, 1).reshape(1, -1, D)), dim=1)
    t = t + pos
    taps = []
    for i in range(L):
        p = f"backbone.encoder.layer.{i}."
        a = p + "attention.attention."
        h = F.layer_norm(t, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
        q = F.linear(h, sd[a + "query.weight"], sd[a + "query.bias"]).view(1, -1, Hh, 64).transpose(1, 2)
        k = F.linear(h, sd[a + "key.weight"], sd[a + "key.bias"]).view(1, -1, Hh, 64).transpose(1, 2)
        v = F.linear(h, sd[a + "value.weight"], sd[a + "value.bias"]).view(1, -1, Hh, 64).transpose(1, 2)
        s = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1)
        o = (s @ v).transpose(1, 2).reshape(1, -1, D)
        o = F.linear(o, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"])
        t = t + o * sd[p + "layer_scale1.lambda1"]
        h = F.layer_norm(t, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
        h = F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
        h = F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        t = t + h * sd[p + "layer_scale2.lambda1"]
        if (i + 1) in cfg["taps"]:
            taps.append(F.layer_norm(t, (D,), sd["backbone.layernorm.weight"], sd["backbone.layernorm.bias"], 1e-6))
    feats = []
    for i, hs in enumerate(taps):
        r = f"neck.reassemble_stage.layers.{i}."
        m = hs[:, 1:].reshape(1, ph, pw, D).permute(0, 3, 1, 2)
        m = F.conv2d(m, sd[r + "projection.weight"], sd[r + "projection.bias"])
        if i == 0:
            m = F.conv_transpose2d(m, sd[r + "resize.weight"], sd[r + "resize.bias"], stride=4)
        elif i == 1:
            m = F.conv_transpose2d(m, sd[r + "resize.weight"], sd[r + "resize.bias"], stride=2)
        elif i == 3:
            m = F.conv2d(m, sd[r + "resize.weight"], sd[r + "resize.bias"], stride=2, padding=1)
        feats.append(F.conv2d(m, sd[f"neck.convs.{i}.weight"], None, padding=1))

    def rl(h, pfx):
        r = h
        h = F.conv2d(F.relu(h), sd[pfx + "convolution1.weight"], sd[pfx + "convolution1.bias"], padding=1)
        h = F.conv2d(F.relu(h), sd[pfx + "convolution2.weight"], sd[pfx + "convolution2.bias"], padding=1)
        return h + r

    fused = None
    rev = feats[::-1]
    fused_all = []
    for j, f in enumerate(rev):
        pfx = f"neck.fusion_stage.layers.{j}."
        h = f if fused is None else fused + rl(f, pfx + "residual_layer1.")
        h = rl(h, pfx + "residual_layer2.")
        if j < 3:
            h = F.interpolate(h, size=rev[j + 1].shape[2:], mode="bilinear", align_corners=True)
        else:
            h = F.interpolate(h, scale_factor=2, mode="bilinear", align_corners=True)
        fused = F.conv2d(h, sd[pfx + "projection.weight"], sd[pfx + "projection.bias"])
        fused_all.append(fused)
    d = F.conv2d(fused, sd["head.conv1.weight"], sd["head.conv1.bias"], padding=1)
    d = F.interpolate(d, (ph * 14, pw * 14), mode="bilinear", align_corners=True)
    d = F.relu(F.conv2d(d, sd["head.conv2.weight"], sd["head.conv2.bias"], padding=1))
    d = F.relu(F.conv2d(d, sd["head.conv3.weight"], sd["head.conv3.bias"]))
    out = d[0, 0]
    if return_parts:
        return out, dict(taps=taps, feats=feats, fused=fused_all, x=t)
    return out
