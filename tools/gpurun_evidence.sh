# Developer tool: the command handed to gpurun for an evidence run; edited per run, outputs under gpurun_out/ (the summaries
# worth keeping are copied to profiles/ by hand).  This version: compute-sanitizer memcheck over the shared-geometry /
# multi-pass / CUDA-graph tests.
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_shared_geometry_gpu.py tests/test_zz_cuda_graph_gpu.py -q > gpurun_out/r1y_sanitizer.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/r1y_sanitizer.log; tail -8 gpurun_out/r1y_sanitizer.log
