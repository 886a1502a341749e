#!/bin/bash
# run 19: pair kernel with 16 epilogue warps: isolation sweep, in-situ spans, pipeline
timeout 300 python tools/gemm_epi_sweep2.py 2>&1 | cut -d'|' -f1-3 | tail -14
for m in 0 3; do
  echo "== spans VD3D_GEMM_2CTA=$m"
  VD3D_GEMM_2CTA=$m timeout 300 python tools/depth_spans.py vitb 4 5 2>&1 | tail -9
  VD3D_GEMM_2CTA=$m timeout 300 python tools/depth_spans.py vitl 4 3 2>&1 | tail -9
done
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"fc1",round(d["roofline"]["frac"],3),"depth_ms",round(d["roofline_depth_stage"]["avg_frame_ms"],3))'
for m in 3; do
  echo "== 1080p VD3D_GEMM_2CTA=$m"; VD3D_GEMM_2CTA=$m timeout 600 python bench.py --no-cpu-baseline --no-4k --steps 10 2>/dev/null | python -c "$P"
  echo "== 4k VD3D_GEMM_2CTA=$m"; VD3D_GEMM_2CTA=$m timeout 600 python bench.py --no-cpu-baseline --workload 4k --steps 5 2>/dev/null | python -c "$P"
done
