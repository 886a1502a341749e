"""Decompose the token GEMMs of the batched forward (DA-V2-Base, 4 frames: M = 10123): launch time by epilogue kind,
kernel variant and debug mode (see GemmArgs.dbg): python tools/gemm_epi_sweep.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from visiondepth3d_b200.depth_engine import DepthEngine  # noqa: E402

e = DepthEngine("vitb", 518, 924)
M = 3 * 2560 + 2443
DBG = {0: "full", 1: "no MMA", 2: "no TMA", 3: "prologue only", 4: "no epilogue", 5: "tmem ld only"}
VAR = {0: "128x128, 3 stages, 2 CTA/SM", 2: "128x128, 3 stages, 1 CTA/SM", 1: "128x128, 6 stages, 1 CTA/SM"}
for name, N, K in (("proj", 768, 768), ("fc2", 768, 3072), ("fc1", 3072, 768), ("qkv", 2304, 768)):
    gf = 2.0 * M * N * K / 1e9
    for epi, act in (("f16", 0), ("resid", 0x100)):
        if epi == "resid" and N != 768:
            continue
        for v in VAR:
            row = []
            for d in DBG:
                ms = e.gemm_bench(M, N, K, variant=v, dbg=d, act=act, iters=20)
                row.append(f"{DBG[d]} {ms * 1e3:6.1f}")
            ms = e.gemm_bench(M, N, K, variant=v, dbg=0, act=act, iters=20)
            print(f"{name:5s} {epi:5s} [{VAR[v]:28s}] {gf / ms:7.1f} TF/s | " + " | ".join(row) + " us", flush=True)
