mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_zz_shared_geometry_gpu.py tests/test_zz_cuda_graph_gpu.py -m gpu -q -x > gpurun_out/r2au_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2au_pytest.log; tail -12 gpurun_out/r2au_pytest.log | cut -c1-200
