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
    # content: what reached the writer equals the frame loop run by hand on the decoded inputs (same decoder, state
    # reset like a fresh render, first pair dropped), bit for bit -- the writer's codec is the only loss
    expected = _by_hand(R, rgb, dep, w, h)
    assert len(expected) == n - 1
    for k, (a, b) in enumerate(zip(frames, expected)):
        mse = float(((a.astype(np.float64) - b) ** 2).mean())
        assert 10 * np.log10(255.0 ** 2 / max(mse, 1e-9)) > 30.0, k   # MJPG round trip of the same picture
    # unsupported options return early like the reference's error paths (no exception escapes, no file left behind)
    out2 = str(tmp_path / "out2.avi")
    assert R.render_sbs_3d(rgb, dep, out2, "MJPG", 24.0, w, h, 4.5, -1.5, -6.0, 0.2, "Half-SBS",
                           _Var("Default (16:9)"), R.aspect_ratios, 0.0, skip_blank_frames=True) is None
    assert not os.path.exists(out2)
    assert R.render_sbs_3d("missing.avi", dep, out2, "MJPG", 24.0, w, h, 4.5, -1.5, -6.0, 0.2, "Half-SBS",
                           _Var("Default (16:9)"), R.aspect_ratios, 0.0) is None
    assert not os.path.exists(out2)


def _by_hand(R, rgb, dep, w, h, **kw):
    import cv2
    rp = R.make_render_params(w, h, 4.5, -1.5, -6.0, 0.2, kw.get("fmt", "Half-SBS"), 16 / 9, 0.0, 10.0, 9, True, True,
                              zero_parallax_strength=0.01, preserve_original_aspect=kw.get("preserve", False))
    R.reset_temporal_state(global_state=False)
    ca, cb = cv2.VideoCapture(rgb), cv2.VideoCapture(dep)
    ca.read(), cb.read()
    outs = []
    while True:
        ok1, f = ca.read()
        ok2, d = cb.read()
        if not (ok1 and ok2):
            break
        outs.append(R.render_frame(f, d, rp))
    return outs


def _write_inputs(tmp_path, w, h, n, kind="natural"):
    import cv2
    from visiondepth3d_b200.synth import synth_frame
    rgb, dep = str(tmp_path / "rgb.avi"), str(tmp_path / "depth.avi")
    fourcc = cv2.VideoWriter_fourcc(*"MJPG")
    wr, wd = cv2.VideoWriter(rgb, fourcc, 24.0, (w, h)), cv2.VideoWriter(dep, fourcc, 24.0, (w, h))
    for i in range(n):
        f, d = synth_frame(i, w, h, kind)
        wr.write(f)
        wd.write(d)
    wr.release()
    wd.release()
    return rgb, dep


def test_render_sbs_3d_ffmpeg_pipe_is_exact(tmp_path, monkeypatch):
    """use_ffmpeg: raw bgr24 frames of out_width x out_height go to the stdin of `ffmpeg ... -f rawvideo -pix_fmt bgr24
    -s WxH -r fps -i - ...` (core/render_3d.py:1143-1163, 1422-1427).  A stand-in `ffmpeg` on PATH records its argv and
    stdin: the bytes must equal the frame loop run by hand, and 20 frames cross three pipelined batches."""
    import stat
    from visiondepth3d_b200 import render_3d as R
    w, h, n = 320, 180, 21
    rgb, dep = _write_inputs(tmp_path, w, h, n)
    fake = tmp_path / "bin"
    fake.mkdir()
    script = fake / "ffmpeg"
    script.write_text('#!/bin/sh\nfor a in "$@"; do last="$a"; done\necho "$@" > "$last.argv"\ncat > "$last.raw"\n')
    script.chmod(script.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(fake) + os.pathsep + os.environ.get("PATH", ""))
    out = str(tmp_path / "out.mp4")
    R.reset_temporal_state()
    R.render_sbs_3d(rgb, dep, out, "mp4v", 24.0, w, h, 4.5, -1.5, -6.0, 0.2, "Full-SBS",
                    _Var("Default (16:9)"), R.aspect_ratios, 0.0, feather_strength=10.0, blur_ksize=9,
                    use_ffmpeg=True, selected_ffmpeg_codec="libx264", crf_value=19,
                    use_subject_tracking=True, use_floating_window=True, preserve_original_aspect=True,
                    suspend_flag=threading.Event(), cancel_flag=threading.Event(), zero_parallax_strength=0.01)
    argv = open(out + ".argv").read().split()
    assert argv[:11] == ["-y", "-f", "rawvideo", "-vcodec", "rawvideo", "-pix_fmt", "bgr24", "-s", f"{2 * w}x{h}", "-r", "24.0"]
    assert "libx264" in argv and argv[argv.index("-crf") + 1] == "19"
    raw = np.fromfile(out + ".raw", dtype=np.uint8)
    assert raw.size == (n - 1) * h * 2 * w * 3
    got = raw.reshape(n - 1, h, 2 * w, 3)
    R.reset_temporal_state()
    expected = _by_hand(R, rgb, dep, w, h, fmt="Full-SBS", preserve=True)
    for k in range(n - 1):
        assert np.array_equal(got[k], expected[k]), k


def test_clip_window_arithmetic():
    """start_s / end_s -> frame indices exactly as core/render_3d.py:1004-1030 computes them."""
    from visiondepth3d_b200.render_3d import _ClipWindow
    wn = _ClipWindow(240, 24.0, None, None)
    assert (wn.empty, wn.first, wn.stop, wn.budget, wn.bounded) == (False, 0, 240, 240, False)
    wn = _ClipWindow(240, 24.0, 1.0, 2.5)
    assert (wn.first, wn.stop, wn.budget, wn.bounded) == (24, 60, 36, True)
    wn = _ClipWindow(240, 23.976, 0.52, 100.0)      # end clamped to the clip, rounding of the start index
    assert (wn.first, wn.stop) == (int(round(0.52 * 23.976)), int(round((240 / 23.976) * 23.976)))
    assert _ClipWindow(240, 24.0, 5.0, 5.0).empty and _ClipWindow(240, 24.0, 11.0, None).empty


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


def test_loader_entry_points_and_depth_video(tmp_path, monkeypatch):
    """ensure_model_downloaded / update_pipeline on a local HF-format checkpoint folder (core/render_depth.py:728-829,
    973-1140), then the depth-video writer in the reference's handoff format (XVID BGR .mkv + .letterbox.json,
    1736-1763, 1894-1935) feeding render_sbs_3d."""
    import cv2
    import json
    import torch
    from safetensors.torch import save_file
    from transformers import DepthAnythingForDepthEstimation
    from visiondepth3d_b200 import render_3d as R
    from visiondepth3d_b200 import render_depth as RD
    from visiondepth3d_b200.depth_weights import hf_config
    # a checkpoint folder laid out like the reference's cache: weights/<org>_<name>/model.safetensors
    torch.manual_seed(0)
    sd = DepthAnythingForDepthEstimation(hf_config("vits")).eval().state_dict()
    wdir = tmp_path / "weights"
    ck = wdir / "depth-anything_Depth-Anything-V2-Small-hf"
    ck.mkdir(parents=True)
    save_file({k: v.contiguous() for k, v in sd.items()}, str(ck / "model.safetensors"))
    monkeypatch.setattr(RD, "local_model_dir", str(wdir))
    got, meta = RD.ensure_model_downloaded("depth-anything/Depth-Anything-V2-Small-hf")
    assert meta["arch"] == "vits" and set(got) == set(sd)
    assert RD.ensure_model_downloaded("depth-anything/Depth-Anything-V2-Base-hf") == (None, None)

    class Label:
        text = ""

        def config(self, text=""):
            Label.text = text

    t = RD.update_pipeline(_Var("Depth Anything V2 Small"), Label(), _Var("Original"), None)
    t.join(timeout=600)
    assert RD.pipe is RD.hf_batch_safe_pipe and RD.pipe_type == "hf" and Label.text.startswith("✅")
    # depth video in the reference's format
    w, h, n = 320, 180, 7
    rgb, _ = _write_inputs(tmp_path, w, h, n)
    dpath = str(tmp_path / "clip_depth.mkv")
    assert RD.depth_video_from_video(rgb, dpath, batch_size=3) == n
    side = json.load(open(str(tmp_path / "clip_depth.letterbox.json")))
    assert side == {"top": 0, "bottom": 0, "orig_w": w, "orig_h": h} == RD.read_letterbox_sidecar(dpath)
    cap = cv2.VideoCapture(dpath)
    assert int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)) == w and int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == n
    ok, d0 = cap.read()
    cap.release()
    assert ok and d0.shape == (h, w, 3) and d0.std() > 1.0
    # ... and it is what render_sbs_3d's depth_path expects
    out = str(tmp_path / "sbs.avi")
    R.render_sbs_3d(rgb, dpath, out, "MJPG", 24.0, w, h, 4.5, -1.5, -6.0, 0.2, "Half-SBS", _Var("Default (16:9)"),
                    R.aspect_ratios, 0.0, feather_strength=10.0, blur_ksize=9, use_subject_tracking=True,
                    use_floating_window=True, suspend_flag=threading.Event(), cancel_flag=threading.Event())
    cap = cv2.VideoCapture(out)
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == n - 1
    cap.release()
    # inference_size path: PIL bicubic in, cv2-INTER_CUBIC-equivalent resize of the u8 depth back out
    d2 = str(tmp_path / "clip_depth_small.mkv")
    assert RD.depth_video_from_video(rgb, d2, inference_size=(224, 126), batch_size=4, max_frames=4) == 4
