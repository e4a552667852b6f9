# round-1 evidence run (gpurun): GPU tests + smoke
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r1w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1w_pytest.log; tail -30 gpurun_out/r1w_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r1w_smoke.log 2>&1; tail -2 gpurun_out/r1w_smoke.log
