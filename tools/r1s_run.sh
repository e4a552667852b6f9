# round-1 evidence run (gpurun): GPU tests, multi-pass timings (two full calls / shared_geometry / forward_passes)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r1x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1x_pytest.log; tail -30 gpurun_out/r1x_pytest.log
timeout 150 python tools/two_pass_times.py --out gpurun_out/r1x_two_pass.json > gpurun_out/r1x_two_pass.log 2>&1; tail -1 gpurun_out/r1x_two_pass.log
timeout 250 python tools/two_pass_times.py --P 4000000 --W 3840 --H 2160 --views 6 --passes 3 --out gpurun_out/r1x_config5_three_pass.json > gpurun_out/r1x_config5.log 2>&1; tail -2 gpurun_out/r1x_config5.log
