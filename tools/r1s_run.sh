# round-1 evidence run (gpurun): GPU tests (incl. CUDA-graph capture), bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r1v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1v_pytest.log; tail -30 gpurun_out/r1v_pytest.log
timeout 200 python bench.py > gpurun_out/r1v_bench_ours.json 2> gpurun_out/r1v_bench_ours.err; head -c 250 gpurun_out/r1v_bench_ours.json; echo
