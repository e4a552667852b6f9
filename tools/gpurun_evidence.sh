mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2ak_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ak_pytest.log; tail -30 gpurun_out/r2ak_pytest.log | cut -c1-250
