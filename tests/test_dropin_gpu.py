"""GPU: the drop-in Python surface itself (video loop of render_sbs_3d, pipe protocol of render_depth)."""
import os
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class _Var:
    def __init__(self, v):
        self.v = v

    def get(self):
        return self.v


def test_render_sbs_3d_video_loop(tmp_path):
    """render_sbs_3d keeps the reference's call surface and sequencing (core/render_3d.py:933-1504):
    n input frames -> n-1 output frames (the first frame of the clip is dropped), Half-SBS geometry."""
    import cv2
    from visiondepth3d_b200 import render_3d as R
    from visiondepth3d_b200.synth import synth_frame
    w, h, n = 320, 180, 6
    rgb, dep, out = str(tmp_path / "rgb.avi"), str(tmp_path / "depth.avi"), str(tmp_path / "out.avi")
    fourcc = cv2.VideoWriter_fourcc(*"MJPG")
    wr, wd = cv2.VideoWriter(rgb, fourcc, 24.0, (w, h)), cv2.VideoWriter(dep, fourcc, 24.0, (w, h))
    assert wr.isOpened() and wd.isOpened()
    for i in range(n):
        f, d = synth_frame(i, w, h, "smooth")
        wr.write(f)
        wd.write(d)
    wr.release()
    wd.release()
    ret = R.render_sbs_3d(rgb, dep, out, "MJPG", 24.0, w, h, 4.5, -1.5, -6.0, 0.2, "Half-SBS",
                          _Var("Default (16:9)"), R.aspect_ratios, 0.0, feather_strength=10.0, blur_ksize=9,
                          use_subject_tracking=True, use_floating_window=True,
                          suspend_flag=threading.Event(), cancel_flag=threading.Event(),
                          zero_parallax_strength=0.01)
    assert ret is None  # the reference returns None
    cap = cv2.VideoCapture(out)
    frames = []
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        frames.append(fr)
    assert len(frames) == n - 1
    assert frames[0].shape == (h, w, 3)
    # the two half-width eyes differ (there is parallax) and are not black
    half = w // 2
    assert frames[2].mean() > 10 and np.abs(frames[2][:, :half].astype(int) - frames[2][:, half:].astype(int)).mean() > 0.05
    # unsupported options return early like the reference's error paths (no exception escapes)
    assert R.render_sbs_3d(rgb, dep, out, "MJPG", 24.0, w, h, 4.5, -1.5, -6.0, 0.2, "Half-SBS",
                           _Var("Default (16:9)"), R.aspect_ratios, 0.0, use_ffmpeg=True) is None
    assert R.render_sbs_3d("missing.avi", dep, out, "MJPG", 24.0, w, h, 4.5, -1.5, -6.0, 0.2, "Half-SBS",
                           _Var("Default (16:9)"), R.aspect_ratios, 0.0) is None


def test_pipe_protocol():
    """pipe(images, inference_size) -> [{"predicted_depth": Tensor[h, w]}] (core/render_depth.py:1113-1119)."""
    import torch
    from PIL import Image
    from visiondepth3d_b200 import render_depth as RD
    from visiondepth3d_b200.synth import synth_frame
    pipe, meta = RD.load_depth_model("vits", width=640, height=360, seed=0)
    assert RD.pipe is pipe and RD.pipe_type == "hf" and meta["processed_size"][0] % 14 == 0
    imgs = [Image.fromarray(synth_frame(i, 640, 360, "smooth")[0][..., ::-1].copy()) for i in range(2)]
    res = pipe(imgs, inference_size=None)
    assert len(res) == 2
    for r_ in res:
        d = r_["predicted_depth"]
        assert isinstance(d, torch.Tensor) and tuple(d.shape) == (360, 640) and torch.isfinite(d).all()
    g = RD.convert_depth_to_grayscale(res[0]["predicted_depth"])
    assert g.dtype == np.uint8 and g.shape == (360, 640) and g.min() == 0 and g.max() >= 250  # +1e-6 in the denominator bites at a 1e-4 range
    fr = synth_frame(0, 640, 360, "smooth")[0]
    assert np.array_equal(RD.depth_u8_from_frame(fr), g)  # GPU min-max u8 == host helper on the same depth
    res2 = pipe([imgs[0]], inference_size=(320, 180))
    assert tuple(res2[0]["predicted_depth"].shape) == (180, 320)
