mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log; tail -3 gpurun_out/r2y_pytest.log
for args in "" "--random --P 300000" "--random --P 1000000 --views 3"; do echo "== new $args"; timeout 300 python tools/stage_times.py $args 2>&1 | grep -E "blend_fwd|blend_bwd|tile_sort|rror"; done
