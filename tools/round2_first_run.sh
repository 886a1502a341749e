# One gpurun call that validates everything written at the end of round 1 without GPU time left:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/round2_first_run.sh'
# 1. the full GPU suite (includes tests/test_fit_gpu.py and the k_post build that carries the opt-in enlarging branch)
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
# 2. the opt-in enlarging eye fit (cv2's fixed-point INTER_AREA emulation); flip the default in plan_fit() if green
VD3D_FIT_ENLARGE=1 timeout 120 python -m pytest tests/test_fit_gpu.py -q -k enlarge 2>&1 | tail -3
# 3. bench sanity of the shipped build
timeout 200 python bench.py --no-cpu-baseline --steps 15 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('1080p', round(d['value'],1), round(d['e2e']['value'],1))"
