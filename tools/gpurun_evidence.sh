mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_zz_reference_callers_gpu.py -m gpu -q -x -k refine_loop > gpurun_out/r2am_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2am_pytest.log; tail -30 gpurun_out/r2am_pytest.log | cut -c1-220
