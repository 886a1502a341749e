# round 2, GPU call 2: ncu of the fast DIBR kernels (1080p + 4K) and the remaining GPU tests
mkdir -p gpurun_out
echo "== pytest gpu (dibr + fit + dropin + abi)"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_depth_gpu.py > gpurun_out/r2_pytest_dibr.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_pytest_dibr.log
echo "== dibr only"; timeout 300 python tools/dibr_only.py 1080p 24; timeout 300 python tools/dibr_only.py 4k 12; timeout 300 python tools/dibr_only.py 1080p 24 --exact
echo "== ncu full 1080p"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_stats|k_shift|k_render" -c 6 -f -o gpurun_out/r2_dibr_1080p python tools/dibr_only.py 1080p 4 --eager > gpurun_out/r2_ncu_1080p.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_1080p.log
echo "== ncu full 4k"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_stats|k_shift|k_render" -c 6 -f -o gpurun_out/r2_dibr_4k python tools/dibr_only.py 4k 4 --eager > gpurun_out/r2_ncu_4k.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_4k.log
ls -la gpurun_out/*.ncu-rep
