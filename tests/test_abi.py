"""CPU: libvd3d.so builds, loads, and exports every symbol include/vd3d.h declares;
ctypes structs mirror the header; no compute is called (no GPU here)."""
import ctypes as C
import os
import re

import pytest

from visiondepth3d_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from visiondepth3d_b200 import build
        build.build()
    return _lib.load()


def _header_functions():
    src = open(os.path.join(ROOT, "include", "vd3d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vd3d_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = _header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in vd3d.h but not exported"
    assert sorted(_lib.SYMBOLS) == sorted(n for n in names)


def test_struct_layouts_match_header(lib):
    # the library reports sizeof() of each ABI struct; the ctypes mirrors must agree
    for which, st in enumerate((_lib.ShiftParams, _lib.RenderParams, _lib.SizePlan, _lib.FrameInfo)):
        assert lib.vd3d_struct_size(which) == C.sizeof(st), st.__name__
    assert C.sizeof(_lib.SizePlan) == 12 * 4
    assert lib.vd3d_struct_size(99) == -1


def test_plan_sizes_matches_oracle(lib):
    from oracle import dibr as O
    from visiondepth3d_b200 import render_3d as R
    cases = [
        (1920, 1080, dict(output_width=1920, output_height=1080, output_format="Half-SBS")),
        (3840, 2160, dict(output_format="Full-SBS", preserve_original_aspect=True)),
        (3840, 2160, dict(output_height=2160, output_format="Full-SBS")),
        (1280, 720, dict(output_height=720, output_format="Red-Cyan Anaglyph")),
        (2048, 858, dict(output_height=858, output_format="Half-SBS", aspect_ratio=2.39)),
        (1440, 1080, dict(output_height=1080, output_format="Half-SBS")),  # 4:3 source, 16:9 target -> crop
        (1920, 800, dict(output_height=1080, output_format="Half-SBS")),   # wide source -> crop width
    ]
    for sw, sh, kw in cases:
        orp = O.RenderParams(**kw)
        rp = R.make_render_params(orp.output_width, orp.output_height, 4.5, -1.5, -6.0, 0.2, orp.output_format,
                                  orp.aspect_ratio, 0.0, preserve_original_aspect=orp.preserve_original_aspect)
        pl = R.plan_sizes(sw, sh, rp)
        op = O.plan_sizes(sw, sh, orp)
        got = tuple(getattr(pl, f) for f, _ in _lib.SizePlan._fields_)
        exp = (op.crop_x0, op.crop_y0, op.crop_w, op.crop_h, op.target_eye_w, op.target_eye_h, op.resized_width,
               op.resized_height, op.per_eye_w, op.per_eye_h, op.out_width, op.out_height)
        assert got == exp, (sw, sh, kw)


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.Vd3dError, match="no CPU fallback"):
        _lib.Context(0)
