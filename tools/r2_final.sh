# round-2 closing run: the whole GPU suite, smoke, the default bench line (both arms + CPU baseline) and the reference arm
mkdir -p gpurun_out
echo "== pytest gpu full"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^FAILED\|passed\|failed" gpurun_out/final_pytest.log | tail -20
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench default"; ( time timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err ) 2>&1 | tail -3; tail -c 300 gpurun_out/final_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
def show(tag,x):
    print(tag,'value',round(x['value'],1),'e2e',round(x['e2e']['value'],1),'fc1',round(x['roofline']['achieved'],1),round(x['roofline']['frac'],3),'depth_ms',round(x['roofline_depth_stage']['avg_frame_ms'],3),round(x['roofline_depth_stage']['frac'],3),'dibr_stage_ms',round(x['roofline_dibr_stage']['avg_frame_ms'],3),round(x['roofline_dibr_stage']['frac'],4),'render_ms',round(x['roofline_dibr_render']['avg_launch_ms'],4),round(x['roofline_dibr_render']['frac'],4),'dibr_only',round(x['dibr_only']['frames_per_s_per_gpu'],1),'py',x['python_surface']['frames_per_s'],x['python_surface']['c_abi_host_buffers_frames_per_s'],'launches',x['gpu_launches'],x['clocks'])
show('1080p',d); show('4k',d['arm_4k']); print('cpu',d.get('cpu_baseline'))
PY
echo "== reference arm (short)"; timeout 600 python bench.py --impl reference --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/final_ref.json; python -c "
import json; d=json.loads(open('gpurun_out/final_ref.json').read()); print(d['value'], d['stage_seconds'], d['wall_s'])"
