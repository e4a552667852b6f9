mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_simple_knn.py -m gpu -q -x > gpurun_out/r2av_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2av_pytest.log; tail -30 gpurun_out/r2av_pytest.log | cut -c1-220
