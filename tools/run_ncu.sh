# ncu --set full captures of the three dominant kernels (one GPU, short command; see B200_PROFILING.md)
set -x
ncu --set full --clock-control none --import-source on -k regex:k_compose -s 3 -c 1 -o gpurun_out/prof_compose python bench.py --workload 4k --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_umma_attention -s 14 -c 1 -o gpurun_out/prof_attn python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_umma_gemm -s 100 -c 6 -o gpurun_out/prof_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out/*.ncu-rep
