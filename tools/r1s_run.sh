mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r1s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1s_pytest.log
tail -15 gpurun_out/r1s_pytest.log
timeout 120 python tools/two_pass_times.py --out gpurun_out/r1s_two_pass.json > gpurun_out/r1s_two_pass.log 2>&1; tail -3 gpurun_out/r1s_two_pass.log
timeout 200 python bench.py > gpurun_out/r1s_bench_ours.json 2> gpurun_out/r1s_bench_ours.err; tail -c 600 gpurun_out/r1s_bench_ours.json
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_shared_geometry_gpu.py -q -k "without_instances or (c_abi and surface_precomp)" > gpurun_out/r1s_sanitizer.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/r1s_sanitizer.log; tail -5 gpurun_out/r1s_sanitizer.log
