mkdir -p gpurun_out
echo "== diag2"; timeout 600 python tools/diag_fast2.py 2>&1 | tail -20
echo "== pytest depth + dropin"; timeout 1500 python -m pytest tests/test_depth_gpu.py tests/test_dropin_gpu.py -m gpu -q --timeout 900 -x > gpurun_out/r4_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^FAILED\|passed\|failed\|Error" gpurun_out/r4_pytest.log | tail -20
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline --steps 6 > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r4_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r4_bench.json').read().strip().splitlines()[-1])
    print('1080p', d['value'], 'e2e', d['e2e']['value'], 'fc1', d['roofline']['achieved'], d['roofline']['frac'], 'depth ms', d['roofline_depth_stage']['avg_frame_ms'], 'dibr stage', d['roofline_dibr_stage']['avg_frame_ms'], d['run'])
    a=d.get('arm_4k')
    if a: print('4k', a['value'], 'e2e', a['e2e']['value'], 'fc1', a['roofline']['achieved'], 'depth ms', a['roofline_depth_stage']['avg_frame_ms'], 'dibr', a['roofline_dibr_stage']['avg_frame_ms'])
except Exception as e:
    print('bench parse failed', e)
PY
for b in 1 2 4; do echo "== batch $b"; VD3D_DEPTH_BATCH=$b timeout 600 python bench.py --no-cpu-baseline --no-4k --steps 4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['achieved'])"; done
