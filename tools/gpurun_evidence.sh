mkdir -p gpurun_out
T=r2bh
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.txt; tail -4 gpurun_out/${T}_pytest.txt | cut -c1-200
timeout 600 python tools/two_pass_times.py --out gpurun_out/${T}_two_pass.json > gpurun_out/${T}_two_pass.log 2>&1; tail -1 gpurun_out/${T}_two_pass.log | cut -c1-900
timeout 600 python tools/two_pass_times.py --P 4000000 --W 3840 --H 2160 --views 6 --passes 3 --out gpurun_out/${T}_config5_three_pass.json > gpurun_out/${T}_three_pass.log 2>&1; tail -1 gpurun_out/${T}_three_pass.log | cut -c1-900
