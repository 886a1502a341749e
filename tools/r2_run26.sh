#!/bin/bash
# run 26: ncu evidence for the round's final kernels (attention v3<2>, pair GEMM, 16-warp residual GEMM) + launch list
mkdir -p gpurun_out
echo "== ncu full, depth batch of 4 (vitb): two transformer layers of the second forward"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_umma|k_layernorm" --launch-skip 250 -c 14 -f -o gpurun_out/r26_depth_vitb python tools/depth_only.py vitb 4 2 > gpurun_out/r26_ncu_depth.log 2>&1; echo "rc=$?"
echo "== ncu launch list of a bench step (1080p)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3600 --csv --log-file gpurun_out/r26_launches_1080p.csv python bench.py --no-4k --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/r26_ncu_bench.log 2>&1; echo "rc=$?"
ls -la gpurun_out/r26*
tail -3 gpurun_out/r26_ncu_depth.log
