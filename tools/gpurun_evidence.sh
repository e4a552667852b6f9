# Developer tool: the command handed to gpurun for an evidence run; edited per run, outputs under gpurun_out/ (the summaries
# worth keeping are copied to profiles/ by hand).  This version: the end-of-round check -- GPU tests, smoke(), both bench arms.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1end_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1end_pytest.log; tail -4 gpurun_out/r1end_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r1end_smoke.log 2>&1; tail -1 gpurun_out/r1end_smoke.log
timeout 120 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r1end_bench_reference.json 2> gpurun_out/r1end_bench_reference.err; head -c 200 gpurun_out/r1end_bench_reference.json; echo
timeout 150 python bench.py > gpurun_out/r1end_bench_ours.json 2> gpurun_out/r1end_bench_ours.err; head -c 200 gpurun_out/r1end_bench_ours.json; echo
