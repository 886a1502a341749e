#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_epi_sweep2.py > gpurun_out/r21_sweep.txt 2>&1
echo "== parity, VD3D_RESID_MODE=1" > gpurun_out/r21_parity.txt
VD3D_RESID_MODE=1 timeout 600 python -m pytest tests/test_depth_gpu.py -q -m gpu 2>&1 | tail -40 >> gpurun_out/r21_parity.txt
echo "== parity, VD3D_RESID_MODE=2 VD3D_GEMM_2CTA=3" >> gpurun_out/r21_parity.txt
VD3D_RESID_MODE=2 VD3D_GEMM_2CTA=3 timeout 600 python -m pytest tests/test_depth_gpu.py -q -m gpu 2>&1 | tail -40 >> gpurun_out/r21_parity.txt
cat gpurun_out/r21_sweep.txt | tail -14
grep -E "FAILED|passed|failed|Error|assert" gpurun_out/r21_parity.txt | head -20
