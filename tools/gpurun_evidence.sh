mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_zz_deterministic_gpu.py -m gpu -q -x > gpurun_out/r2ax_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ax_pytest.log; tail -40 gpurun_out/r2ax_pytest.log | cut -c1-220
