#!/bin/bash
# run 23: attention kernel generations: parity of the forward + in-situ device time of the attention launches
mkdir -p gpurun_out
for m in -1 0 1 2; do
  echo "== VD3D_ATTN_MODE=$m"
  VD3D_ATTN_MODE=$m timeout 600 python -m pytest tests/test_depth_gpu.py -q -m gpu -k "forward_matches_oracle or infer_batch_equals or outlier" 2>&1 | tail -12 > gpurun_out/r23_pytest_$m.txt
  grep -E "passed|failed|Error|assert " gpurun_out/r23_pytest_$m.txt | head -6
  VD3D_ATTN_MODE=$m timeout 300 python tools/depth_spans.py vitb 4 5 2>&1 | grep -E "attn|sum"
  VD3D_ATTN_MODE=$m timeout 300 python tools/depth_spans.py vitl 4 3 2>&1 | grep -E "attn|sum"
done
