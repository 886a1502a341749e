# depth kernel parity + both bench workloads (no CPU baseline): quick A/B after a kernel change
timeout 600 python -m pytest tests/test_depth_gpu.py -q 2>&1 | tail -3
for w in 1080p 4k; do
  st=15; [ $w = 4k ] && st=6
  timeout 300 python bench.py --workload $w --no-cpu-baseline --steps $st --warmup 3 > gpurun_out/check_$w.json 2> gpurun_out/check_$w.err
  python -c "
import json;d=json.loads(open('gpurun_out/check_$w.json').read().strip().splitlines()[-1]);print('$w', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['achieved'],1), round(d['roofline_depth_stage']['avg_frame_ms'],3), round(d['roofline_dibr_stage']['avg_frame_ms'],3))"
done
