"""CTA-pair variants of the token GEMMs at the batched row count (DA-V2-Base, 4 frames: M = 10123; Large: D = 1024)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from visiondepth3d_b200.depth_engine import DepthEngine  # noqa: E402

e = DepthEngine("vitb", 518, 924)
M = 3 * 2560 + 2443
VAR = {0: "1-CTA 128x128 s3 x2/SM", 1: "1-CTA s6 x1/SM", 3: "1-CTA s6 x1/SM 16 epi warps", 10: "pair 256x256 s6"}
for name, N, K in (("proj", 768, 768), ("fc2", 768, 3072), ("fc1", 3072, 768), ("qkv", 2304, 768), ("fc1-L", 4096, 1024),
                   ("fc2-L", 1024, 4096), ("qkv-L", 3072, 1024), ("proj-L", 1024, 1024)):
    gf = 2.0 * M * N * K / 1e9
    for epi, act in (("f16", 0), ("resid", 0x100)):
        if epi == "resid" and N > 1024:
            continue
        row = []
        for v in VAR:
            ms = e.gemm_bench(M, N, K, variant=v, dbg=0, act=act, iters=20)
            row.append(f"{VAR[v]} {ms * 1e3:6.1f} us {gf / ms:6.0f} TF/s")
        print(f"{name:6s} {epi:5s} | " + " | ".join(row), flush=True)
