# final-build evidence: launch list of the bench command + --set full of the GEMM (6 launches) and attention kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_1080p_final3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_umma_gemm -s 100 -c 6 -o gpurun_out/prof_gemm_final3 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_umma_attention -s 14 -c 1 -o gpurun_out/prof_attn_final3 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/*final3*
