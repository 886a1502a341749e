# round 2, GPU call 1: first run of the fast DIBR path (k_stats / k_render), the enlarging eye fit, the graph-epoch fix
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r1_gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r1_smoke.log
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -x --deselect tests/test_depth_gpu.py > gpurun_out/r1_pytest_dibr.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r1_pytest_dibr.log
echo "== pytest depth"; timeout 900 python -m pytest tests/test_depth_gpu.py -m gpu -q --timeout 600 > gpurun_out/r1_pytest_depth.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r1_pytest_depth.log
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline --steps 6 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r1_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r1_bench.json').read().strip().splitlines()[-1])
    for k in ('value',):
        print('1080p', d['value'], 'e2e', d['e2e']['value'], 'dibr_only', d['dibr_only'], 'stage', d['roofline_dibr_stage']['avg_frame_ms'], 'render', d['roofline_dibr_render']['avg_launch_ms'], d['run'].get('graphs_active'))
    a=d.get('arm_4k')
    if a: print('4k', a['value'], 'e2e', a['e2e']['value'], 'dibr_only', a['dibr_only'], 'stage', a['roofline_dibr_stage']['avg_frame_ms'], 'render', a['roofline_dibr_render']['avg_launch_ms'])
except Exception as e:
    print('bench parse failed', e)
PY
