mkdir -p gpurun_out
nvidia-smi -L
echo "== 2-GPU exact sharding test"; timeout 900 python -m pytest tests/test_sharding_gpu.py -m gpu -q --timeout 600 2>&1 | tail -5
echo "== bench N=2 replicas"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 --no-4k > gpurun_out/r7_bench_n2.json 2> gpurun_out/r7_bench_n2.err; echo "rc=$?"; tail -c 400 gpurun_out/r7_bench_n2.err; python -c "
import json; d=json.loads(open('gpurun_out/r7_bench_n2.json').read().strip().splitlines()[-1]); print('N=2 replicas value', d['value'], 'e2e', d['e2e']['value'])"
echo "== bench N=2 exact"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 10 --warmup 3 --sharding exact > gpurun_out/r7_bench_n2_exact.json 2> gpurun_out/r7_bench_n2_exact.err; echo "rc=$?"; tail -c 600 gpurun_out/r7_bench_n2_exact.err; tail -1 gpurun_out/r7_bench_n2_exact.json | cut -c1-1500
echo "== bench N=1 exact (same code path, one rank)"; timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --sharding exact 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1 exact-path value', d['value'], d['run'])"
echo "== bench N=1 default (reference for the ratio)"; timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-4k --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1 value', d['value'], 'e2e', d['e2e']['value'], 'fc1', d['roofline']['achieved'], 'py', d['python_surface']['frames_per_s'], d['python_surface']['c_abi_host_buffers_frames_per_s'])"
