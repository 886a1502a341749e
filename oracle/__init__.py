"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements (numpy / torch-fp32) of the reference algorithms on the
depth -> stereo hot path of VisionDepth3D (core/render_3d.py,
core/render_depth.py + HF transformers Depth-Anything-V2).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this package, and only as the checker or
as the timed CPU baseline.  The product (visiondepth3d_b200/) never imports
it and has no CPU fallback: it fails loudly when libvd3d.so is missing.

Pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the UNMODIFIED reference run in the
build container through tools/refshim.py; the vectors live in tests/golden/
with the generating script tools/gen_golden.py.
"""
