mkdir -p gpurun_out
echo "== pytest gpu full"; timeout 3000 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r5_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "^FAILED\|passed\|failed" gpurun_out/r5_pytest.log | tail -30
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== dibr only"; timeout 300 python tools/dibr_only.py 1080p 24 | tail -1; timeout 300 python tools/dibr_only.py 4k 12 | tail -1
echo "== ncu launch list of a bench step (1080p)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r5_launches_1080p.csv python bench.py --no-4k --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/r5_ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full DIBR 1080p / 4k"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_stats|k_shift|k_render" -c 3 -f -o gpurun_out/r5_dibr_1080p python tools/dibr_only.py 1080p 4 --eager > gpurun_out/r5_ncu_1080p.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_stats|k_shift|k_render" -c 3 -f -o gpurun_out/r5_dibr_4k python tools/dibr_only.py 4k 4 --eager > gpurun_out/r5_ncu_4k.log 2>&1; echo "rc=$?"
echo "== ncu full depth batch (vitb, B=3): first transformer layers"
timeout 900 ncu --set full --clock-control none -k regex:"k_umma|k_layernorm" --launch-skip 6 -c 10 -f -o gpurun_out/r5_depth_vitb python tools/depth_only.py vitb 3 1 > gpurun_out/r5_ncu_depth.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r5_launches_1080p.csv
