"""CPU: the host pipeline of the drop-in render_sbs_3d (reader workers -> pinned ring -> batched render calls -> writer
thread) against a stub of the C ABI: ordering, frame counts, first pair dropped, clip window, cancel, early refusals.
No GPU and no oracle involved: this is host logic (tests/test_dropin_gpu.py covers the real library)."""
import ctypes as C
import threading

import numpy as np
import pytest


class _Var:
    def __init__(self, v):
        self.v = v

    def get(self):
        return self.v


class _StubLib:
    """vd3d_render_clip 'renders' by writing each input frame's first byte into the output: enough to check ordering."""

    def __init__(self, real):
        self.real = real
        self.calls = []
        self._keep = []

    def vd3d_plan_sizes(self, *a):
        return self.real.vd3d_plan_sizes(*a)

    def vd3d_check_config(self, *a):
        return 0

    def vd3d_sync(self, *a):
        return 0

    def vd3d_host_alloc(self, n):
        buf = (C.c_uint8 * int(n))()
        self._keep.append(buf)
        return C.addressof(buf)

    def vd3d_host_free(self, p):
        return None

    def vd3d_render_clip(self, h, n, fp, dp, dch, sh, sw, rp, op, mem, infos):
        self.calls.append(n)
        for i in range(n):
            tag = C.cast(fp[i], C.POINTER(C.c_uint8))[0]
            dtag = C.cast(dp[i], C.POINTER(C.c_uint8))[0]
            out = C.cast(op[i], C.POINTER(C.c_uint8))
            out[0] = tag
            out[1] = dtag
        return 0


class _StubCtx:
    def __init__(self, lib):
        self.lib, self.h = lib, C.c_void_p(1)

    def check(self, rc):
        assert rc == 0

    def reset(self, which=3):
        pass


@pytest.fixture
def loop(monkeypatch):
    import cv2
    from visiondepth3d_b200 import _lib
    from visiondepth3d_b200 import render_3d as R
    lib = _StubLib(_lib.load())
    monkeypatch.setattr(R, "_ctx", lambda: _StubCtx(lib))
    state = {"n": 30, "fps": 24.0}

    class Cap:
        def __init__(self, path):
            self.depth, self.pos, self.ok = path == "depth", 0, path != "missing"

        def isOpened(self):
            return self.ok

        def get(self, prop):
            return {cv2.CAP_PROP_FRAME_COUNT: float(state["n"]), cv2.CAP_PROP_FPS: state["fps"],
                    cv2.CAP_PROP_POS_FRAMES: float(self.pos)}.get(prop, 0.0)

        def set(self, prop, v):
            if prop == cv2.CAP_PROP_POS_FRAMES:
                self.pos = int(v)
            return True

        def read(self):
            if self.pos >= (state.get("n_depth", state["n"]) if self.depth else state["n"]):
                return False, None
            f = np.full((36, 64, 3), (self.pos + (100 if self.depth else 0)) % 256, dtype=np.uint8)
            self.pos += 1
            return True, f

        def release(self):
            pass

    written = []

    class Sink:
        def __init__(self, path, size, *a):
            self.size = size

        def ok(self):
            return True

        def write(self, frame):
            written.append((int(frame.flat[0]), int(frame.flat[1]), frame.shape))

        def close(self):
            pass

    monkeypatch.setattr(cv2, "VideoCapture", Cap)
    monkeypatch.setattr(R, "_FrameSink", Sink)

    def run(**kw):
        written.clear()
        lib.calls.clear()
        R.render_sbs_3d(kw.pop("src", "rgb"), "depth", "out.mp4", "mp4v", 24.0, 64, 36, 4.5, -1.5, -6.0, 0.2, "Half-SBS",
                        _Var("Default (16:9)"), R.aspect_ratios, 0.0, suspend_flag=threading.Event(),
                        cancel_flag=kw.pop("cancel", threading.Event()), **kw)
        return list(written), list(lib.calls)
    return run, state


def test_pipeline_order_counts_and_batches(loop):
    run, state = loop
    out, calls = run()
    # 30 frames in the clip: the first pair is dropped, pairs stay paired, order is kept, batches of 8
    assert [o[0] for o in out] == list(range(1, 30)) and [o[1] for o in out] == list(range(101, 130))
    assert out[0][2] == (36, 64, 3) and calls == [8, 8, 8, 5]
    state["n"] = 9
    out, calls = run()
    assert [o[0] for o in out] == list(range(1, 9)) and calls == [8]
    state["n"] = 1          # nothing left after the dropped pair
    assert run() == ([], [])


def test_clip_window_and_refusals(loop):
    run, state = loop
    state["n"] = 240
    out, _ = run(start_s=1.0, end_s=2.0)          # frames 24 .. 47: first pair (24) dropped, stops at the window's end
    assert [o[0] for o in out] == list(range(25, 48))
    assert run(start_s=5.0, end_s=5.0) == ([], [])            # empty window
    assert run(skip_blank_frames=True) == ([], [])            # refused before anything is opened
    assert run(src="missing") == ([], [])
    ev = threading.Event()
    ev.set()
    assert run(cancel=ev)[0] == []                            # cancelled before the first frame


def test_lookahead_keeps_order_and_stops_at_the_shorter_stream(loop):
    """The stream workers run several pairs ahead of the published position: order, pairing and the end of the clip must
    not depend on it."""
    run, state = loop
    state["n"] = 100
    out, calls = run()
    assert [o[0] for o in out] == list(range(1, 100)) and [o[1] for o in out] == [(100 + i) % 256 for i in range(1, 100)]
    assert calls == [8] * 12 + [3]
    state["n_depth"] = 41                          # depth video shorter than the colour video: pairs end with it
    out, _ = run()
    assert [o[0] for o in out] == list(range(1, 41))
    del state["n_depth"]
    state["n"] = 240
    out, _ = run(start_s=8.0)                      # open-ended window: from frame 192 (dropped) to the end of the clip
    assert [o[0] for o in out] == list(range(193, 240))
    out, _ = run(start_s=2.0, end_s=2.05)          # one-frame window: the pair after the dropped one is still rendered
    assert [o[0] for o in out] == [49]
