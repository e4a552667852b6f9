mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_fused_passes_gpu.py -m gpu -q -x > gpurun_out/r2br_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2br_pytest.log; tail -25 gpurun_out/r2br_pytest.log | cut -c1-250
