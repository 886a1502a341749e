#!/bin/bash
# run 24: attention v3<2> as default; depth batch size against wave quantisation
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"fc1",round(d["roofline"]["frac"],3),"depth_ms",round(d["roofline_depth_stage"]["avg_frame_ms"],3))'
for b in 4 5 8 3; do
  echo "== 1080p depth batch $b"; VD3D_DEPTH_BATCH=$b timeout 600 python bench.py --no-cpu-baseline --no-4k --steps 12 2>/dev/null | python -c "$P"
done
for b in 4 6 3; do
  echo "== 4k depth batch $b"; VD3D_DEPTH_BATCH=$b timeout 600 python bench.py --no-cpu-baseline --workload 4k --steps 6 2>/dev/null | python -c "$P"
done
