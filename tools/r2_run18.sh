#!/bin/bash
# run 18: pair-GEMM policy (VD3D_GEMM_2CTA=3) inside the pipeline + parity of the forward with it
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"fc1",round(d["roofline"]["frac"],3),"depth_ms",round(d["roofline_depth_stage"]["avg_frame_ms"],3))'
for m in 0 3; do
  echo "== 1080p VD3D_GEMM_2CTA=$m"; VD3D_GEMM_2CTA=$m timeout 600 python bench.py --no-cpu-baseline --no-4k --steps 10 2>/dev/null | python -c "$P"
  echo "== 4k VD3D_GEMM_2CTA=$m"; VD3D_GEMM_2CTA=$m timeout 600 python bench.py --no-cpu-baseline --workload 4k --steps 5 2>/dev/null | python -c "$P"
done
echo "== depth tests with the pair policy"
VD3D_GEMM_2CTA=3 timeout 900 python -m pytest tests/test_depth_gpu.py -x -q -m gpu 2>&1 | tail -4
