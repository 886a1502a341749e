mkdir -p gpurun_out
echo "== dropin + sr tests"; timeout 900 python -m pytest tests/test_dropin_gpu.py tests/test_sr_gpu.py tests/test_fit_gpu.py -m gpu -q --timeout 300 2>&1 | tail -12
echo "== sr timing"; timeout 300 python tools/sr_only.py 1920 1080 32 6 2>&1 | tail -1; timeout 300 python tools/sr_only.py 960 540 32 8 2>&1 | tail -1
echo "== bench default"; ( time timeout 900 python bench.py > gpurun_out/r9_bench.json 2> gpurun_out/r9_bench.err ) 2>&1 | tail -3; tail -c 300 gpurun_out/r9_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r9_bench.json').read().strip().splitlines()[-1])
def show(tag,x):
    print(tag,'value',round(x['value'],1),'e2e',round(x['e2e']['value'],1),'fc1',round(x['roofline']['achieved'],1),round(x['roofline']['frac'],3),'depth_ms',round(x['roofline_depth_stage']['avg_frame_ms'],3),round(x['roofline_depth_stage']['frac'],3),'dibr_stage_ms',round(x['roofline_dibr_stage']['avg_frame_ms'],3),round(x['roofline_dibr_stage']['frac'],4),'render_ms',round(x['roofline_dibr_render']['avg_launch_ms'],4),round(x['roofline_dibr_render']['frac'],4),'dibr_only',round(x['dibr_only']['frames_per_s_per_gpu'],1),round(x['dibr_only']['frac'],4),'py',x['python_surface']['frames_per_s'],x['python_surface']['c_abi_host_buffers_frames_per_s'],'launches',x['gpu_launches'],x['clocks'])
show('1080p',d); show('4k',d['arm_4k']); print('cpu',d.get('cpu_baseline'))
PY
