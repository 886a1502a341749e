#!/bin/bash
# run 25: polynomial exp2 share in attention v3<2>; paired pipeline comparison old / new attention on one box
mkdir -p gpurun_out
for m in 2 3 4; do
  echo "== VD3D_ATTN_MODE=$m"
  VD3D_ATTN_MODE=$m timeout 300 python tools/depth_spans.py vitb 4 5 2>&1 | grep -E "attn|sum"
  VD3D_ATTN_MODE=$m timeout 300 python tools/depth_spans.py vitl 4 3 2>&1 | grep -E "attn|sum"
done
for m in 3 4; do
  VD3D_ATTN_MODE=$m timeout 600 python -m pytest tests/test_depth_gpu.py -q -m gpu -k "forward_matches_oracle or outlier" 2>&1 | tail -12 > gpurun_out/r25_pytest_$m.txt
  echo "parity mode $m:"; grep -E "passed|failed|Error|assert " gpurun_out/r25_pytest_$m.txt | head -6
done
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"fc1",round(d["roofline"]["frac"],3),"depth_ms",round(d["roofline_depth_stage"]["avg_frame_ms"],3))'
for m in -1 2 3 -1 2 3; do
  echo "== 1080p VD3D_ATTN_MODE=$m"; VD3D_ATTN_MODE=$m timeout 600 python bench.py --no-cpu-baseline --no-4k --steps 12 2>/dev/null | python -c "$P"
done
