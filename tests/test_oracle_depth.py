"""CPU: pin oracle/depth.py against the installed transformers DepthAnythingForDepthEstimation
(the third-party module the reference calls; random-init weights, seed 0)."""
import torch

from oracle import depth as OD
from visiondepth3d_b200.depth_weights import CONFIGS, hf_config


def test_oracle_matches_transformers_small_input():
    from transformers import DepthAnythingForDepthEstimation
    torch.manual_seed(0)
    model = DepthAnythingForDepthEstimation(hf_config("vits")).eval()
    sd = model.state_dict()
    torch.manual_seed(1)
    for shape in ((3, 70, 98), (3, 112, 154)):
        px = torch.randn(*shape)
        with torch.no_grad():
            ref = model(pixel_values=px[None]).predicted_depth[0]
            mine = OD.forward(sd, CONFIGS["vits"], px)
        assert mine.shape == ref.shape
        scale = float(ref.max() - ref.min()) + 1e-12
        assert float((mine - ref).abs().max()) / scale < 1e-4


import pytest


@pytest.mark.parametrize("name", ["vitb", "vitl"])
def test_oracle_matches_transformers_base_and_large_configs(name):
    """The Base / Large configurations (taps, neck widths, fusion width, 24 layers) through the same oracle code:
    the GPU tests compare the CUDA engine with oracle/depth.py for these, so the oracle itself must equal transformers."""
    from transformers import DepthAnythingForDepthEstimation
    torch.manual_seed(0)
    model = DepthAnythingForDepthEstimation(hf_config(name)).eval()
    sd = model.state_dict()
    torch.manual_seed(2)
    px = torch.randn(3, 70, 98)
    with torch.no_grad():
        ref = model(pixel_values=px[None]).predicted_depth[0]
        mine = OD.forward(sd, CONFIGS[name], px)
    assert mine.shape == ref.shape
    scale = float(ref.max() - ref.min()) + 1e-12
    assert float((mine - ref).abs().max()) / scale < 1e-4


def test_weight_preparation_shapes():
    from transformers import DepthAnythingForDepthEstimation
    from visiondepth3d_b200.depth_weights import prepare
    torch.manual_seed(0)
    sd = DepthAnythingForDepthEstimation(hf_config("vits")).state_dict()
    w = prepare(sd, CONFIGS["vits"], 518, 924)
    assert w["pe.w"].shape == (384, 592) and w["pos"].shape == (37 * 66 + 1, 384)
    assert w["l0.qkv.w"].shape == (1152, 384) and w["l0.qkv.w"].dtype.name == "float16"
    assert w["r0.proj.w"].shape == (64, 384) and w["r0.up.w"].shape == (16 * 64, 64)
    assert w["n0.conv.w"].shape == (64, 9 * 64) and w["r3.down.w"].shape == (384, 9 * 384)
    assert w["h.c1.w"].shape == (64, 9 * 64) and w["h.c2.w"].shape == (32, 9 * 64) and w["h.c3.w"].shape == (32,)
