# Developer tool: the command handed to gpurun for an evidence run; edited per run, outputs under gpurun_out/ (the summaries
# worth keeping are copied to profiles/ by hand).  This version: the tests that exercise k_recolor (re-blend, graph capture).
mkdir -p gpurun_out
python -m pytest tests/test_zz_cuda_graph_gpu.py tests/test_zz_shared_geometry_gpu.py -q -x > gpurun_out/r1last_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1last_pytest.log; tail -5 gpurun_out/r1last_pytest.log
