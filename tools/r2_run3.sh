mkdir -p gpurun_out
echo "== diag"; timeout 600 python tools/diag_fast.py 2>&1 | tail -40
echo "== dibr only"; timeout 300 python tools/dibr_only.py 1080p 24 | tail -1; timeout 300 python tools/dibr_only.py 4k 12 | tail -1
echo "== pytest"; timeout 2400 python -m pytest tests/test_dibr_gpu.py tests/test_fit_gpu.py tests/test_dropin_gpu.py -m gpu -q --timeout 900 -k "not vs_oracle[" > gpurun_out/r3_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^FAILED\|passed\|failed" gpurun_out/r3_pytest.log | tail -20
