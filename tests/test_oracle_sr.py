"""CPU: the pre / post-processing chain of the Real-ESRGAN stage (core/merged_pipeline.py:221-267) restated in
oracle/sr.py against the real cv2 ops, and structural checks of the SRVGGNetCompact restatement (the network itself is
third-party and absent offline: parity unpinned, see oracle/sr.py)."""
import numpy as np

from oracle import sr as S


def test_pre_post_blend_match_cv2():
    import cv2
    rng = np.random.default_rng(3)
    fr = rng.integers(0, 256, (36, 64, 3), dtype=np.uint8)
    ref = (cv2.cvtColor(fr, cv2.COLOR_BGR2RGB).astype(np.float32) / 255.0).transpose(2, 0, 1)[None]
    assert np.array_equal(S.preprocess_esr(fr), ref)
    t = rng.random((1, 3, 36, 64)).astype(np.float32) * 1.4 - 0.2
    ref = cv2.cvtColor((np.clip(t[0].transpose(1, 2, 0), 0, 1) * 255.0).astype(np.uint8), cv2.COLOR_RGB2BGR)
    assert np.array_equal(S.postprocess_esr(t), ref)
    up = rng.integers(0, 256, (36, 64, 3), dtype=np.uint8)
    for mode, a in (("LOW", 0.85), ("MEDIUM", 0.5), ("HIGH", 0.25)):
        assert np.array_equal(S.blend_images(fr, up, mode), cv2.addWeighted(up, a, fr, 1 - a, 0)), mode
    assert S.blend_images(fr, up, "OFF") is up


def test_cubic_bgr_matches_cv2():
    import cv2
    rng = np.random.default_rng(4)
    img = cv2.GaussianBlur(rng.integers(0, 256, (144, 256, 3), dtype=np.uint8), (0, 0), 1.5)
    for (ow, oh) in ((64, 36), (100, 57), (512, 288), (256, 144)):
        d = np.abs(S.resize_cubic_bgr(img, ow, oh).astype(int) - cv2.resize(img, (ow, oh), interpolation=cv2.INTER_CUBIC).astype(int))
        assert d.max() <= 1 and (d > 0).mean() <= 2e-3, (ow, oh, d.max(), (d > 0).mean())


def test_srvgg_structure():
    """Shapes, the pixel-shuffle channel order and the nearest-upsampled base of SRVGGNetCompact (srvgg_arch.py)."""
    import torch
    sd = S.srvgg_state_dict(num_conv=4, seed=1)
    assert sd["body.0.weight"].shape == (64, 3, 3, 3) and sd["body.10.weight"].shape == (48, 64, 3, 3)
    x = torch.rand(1, 3, 9, 13)
    y = S.srvgg_forward(sd, x)
    assert tuple(y.shape) == (1, 3, 36, 52)
    zero = {k: (torch.zeros_like(v) if "weight" in k and v.dim() == 4 or k.endswith("bias") else v) for k, v in sd.items()}
    y0 = S.srvgg_forward(zero, x)                      # all-zero convs: the output is the nearest x4 of the input
    assert torch.equal(y0, x.repeat_interleave(4, 2).repeat_interleave(4, 3))
    zero["body.10.bias"] = torch.arange(48, dtype=torch.float32)
    y1 = S.srvgg_forward(zero, torch.zeros(1, 3, 2, 2))
    assert float(y1[0, 1, 2, 3]) == 16 + 2 * 4 + 3      # out[c, 4h+i, 4w+j] = conv[c*16 + i*4 + j, h, w]
