for m in 1 2; do
  echo "== mode $m tests"; VD3D_GEMM_2CTA=$m timeout 400 python -m pytest tests/test_depth_gpu.py -q 2>&1 | tail -4
done
for m in 0 1 2; do
  VD3D_GEMM_2CTA=$m timeout 200 python bench.py --no-cpu-baseline --steps 15 --warmup 3 > gpurun_out/bench_pair${m}_1080p.json 2> gpurun_out/bench_pair${m}_1080p.err
  python -c "
import json;d=json.loads(open('gpurun_out/bench_pair${m}_1080p.json').read().strip().splitlines()[-1]);print('1080p mode $m', d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline_depth_stage']['avg_frame_ms'])"
done
for m in 0 2; do
  VD3D_GEMM_2CTA=$m timeout 200 python bench.py --workload 4k --no-cpu-baseline --steps 6 --warmup 3 > gpurun_out/bench_pair${m}_4k.json 2> gpurun_out/bench_pair${m}_4k.err
  python -c "
import json;d=json.loads(open('gpurun_out/bench_pair${m}_4k.json').read().strip().splitlines()[-1]);print('4k mode $m', d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline_depth_stage']['avg_frame_ms'])"
done
