"""Sweep the tcgen05 GEMM kernel variants on the transformer shapes of DA-V2 (tuning hook vd3d_gemm_bench).
dbg: 1 operand feed only, 2 MMAs only, 3 prologue+teardown, 4 no epilogue work, 5 epilogue reads TMEM only; +8 spin waits."""
import sys
sys.path.insert(0, ".")
from visiondepth3d_b200.depth_engine import DepthEngine

eng = DepthEngine("vits", 70, 98)
SHAPES = {"qkv_b": (2443, 2304, 768), "proj_b": (2443, 768, 768), "fc1_b": (2443, 3072, 768), "fc2_b": (2443, 768, 3072),
          "fc1_l": (2443, 4096, 1024), "big": (8192, 8192, 1024)}
VARIANTS = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 1, 10]
DBG = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 4, 5, 8, 9, 10, 12, 13]
print("shape        var dbg   us     TFLOP/s")
for name, (M, N, K) in SHAPES.items():
    for v in VARIANTS:
        for dbg in DBG:
            try:
                ms = eng.gemm_bench(M, N, K, v, dbg, 0, 20)
            except Exception as ex:  # noqa: BLE001
                print(f"{name:12s} {v:3d} {dbg:3d}   error {ex}")
                continue
            print(f"{name:12s} {v:3d} {dbg:3d} {ms * 1e3:7.1f} {2.0 * M * N * K / (ms * 1e-3) / 1e12:8.1f}", flush=True)
