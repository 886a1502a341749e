mkdir -p gpurun_out
echo "== pytest dibr + sr + fit"; timeout 1500 python -m pytest tests/test_dibr_gpu.py tests/test_sr_gpu.py tests/test_fit_gpu.py -m gpu -q --timeout 600 > gpurun_out/r11_pytest.log 2>&1; echo "rc=$?"; grep -n "^FAILED\|passed\|failed" gpurun_out/r11_pytest.log | tail -12
echo "== dibr only"; timeout 300 python tools/dibr_only.py 1080p 24 | tail -1; timeout 300 python tools/dibr_only.py 4k 12 | tail -1
echo "== ncu times"; timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_stats|k_shift|k_render" -c 3 python tools/dibr_only.py 1080p 4 --eager 2>&1 | grep -E "k_stats|k_shift|k_render|duration|inst_executed" | head -12
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_stats|k_shift|k_render" -c 3 python tools/dibr_only.py 4k 4 --eager 2>&1 | grep -E "k_stats|k_shift|k_render|duration|inst_executed" | head -12
