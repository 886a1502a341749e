#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_depth_gpu.py tests/test_sr_gpu.py tests/test_dropin_gpu.py -q -m gpu 2>&1 | tail -15 > gpurun_out/r22_pytest.txt
tail -5 gpurun_out/r22_pytest.txt
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"fc1",round(d["roofline"]["frac"],3),"depth_ms",round(d["roofline_depth_stage"]["avg_frame_ms"],3))'
echo "== 1080p"; timeout 600 python bench.py --no-cpu-baseline --no-4k --steps 20 2>/dev/null | python -c "$P"
echo "== 4k"; timeout 600 python bench.py --no-cpu-baseline --workload 4k --steps 10 2>/dev/null | python -c "$P"
echo "== 1080p policy off"; VD3D_GEMM_POLICY=0 timeout 600 python bench.py --no-cpu-baseline --no-4k --steps 20 2>/dev/null | python -c "$P"
echo "== 4k policy off"; VD3D_GEMM_POLICY=0 timeout 600 python bench.py --no-cpu-baseline --workload 4k --steps 10 2>/dev/null | python -c "$P"
