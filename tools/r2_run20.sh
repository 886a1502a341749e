#!/bin/bash
# run 20: residual-stream prefetch (1-CTA 16 epilogue warps; pair kernel), isolation + parity + in-situ
timeout 300 python tools/gemm_epi_sweep2.py 2>&1 | tail -14
echo "== parity, VD3D_RESID_MODE=1"
VD3D_RESID_MODE=1 timeout 600 python -m pytest tests/test_depth_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== parity, VD3D_RESID_MODE=2 VD3D_GEMM_2CTA=3"
VD3D_RESID_MODE=2 VD3D_GEMM_2CTA=3 timeout 600 python -m pytest tests/test_depth_gpu.py -x -q -m gpu 2>&1 | tail -3
for cfg in "0 0" "1 0" "2 0" "2 3"; do
  set -- $cfg
  echo "== spans VD3D_RESID_MODE=$1 VD3D_GEMM_2CTA=$2"
  VD3D_RESID_MODE=$1 VD3D_GEMM_2CTA=$2 timeout 300 python tools/depth_spans.py vitb 4 5 2>&1 | grep -E "qkv|proj|fc1|fc2|sum"
  VD3D_RESID_MODE=$1 VD3D_GEMM_2CTA=$2 timeout 300 python tools/depth_spans.py vitl 4 3 2>&1 | grep -E "qkv|proj|fc1|fc2|sum"
done
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"fc1",round(d["roofline"]["frac"],3),"depth_ms",round(d["roofline_depth_stage"]["avg_frame_ms"],3))'
for cfg in "1 0" "2 0" "2 3"; do
  set -- $cfg
  echo "== 1080p VD3D_RESID_MODE=$1 VD3D_GEMM_2CTA=$2"; VD3D_RESID_MODE=$1 VD3D_GEMM_2CTA=$2 timeout 600 python bench.py --no-cpu-baseline --no-4k --steps 10 2>/dev/null | python -c "$P"
  echo "== 4k VD3D_RESID_MODE=$1 VD3D_GEMM_2CTA=$2"; VD3D_RESID_MODE=$1 VD3D_GEMM_2CTA=$2 timeout 600 python bench.py --no-cpu-baseline --workload 4k --steps 5 2>/dev/null | python -c "$P"
done
