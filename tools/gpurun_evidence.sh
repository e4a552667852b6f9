# Developer tool: the command handed to gpurun for an evidence run; edited per run, outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log; tail -4 gpurun_out/r2g_pytest.log
timeout 300 python tools/stage_times.py 2>&1 | grep -E "blend_fwd|tile_sort|blend_bwd"
ncu --set full --clock-control none --import-source on -k regex:k_blend_fwd -s 2 -c 1 -o gpurun_out/r2g_fwd python tools/profile_view.py --views 3 > gpurun_out/r2g_fwd.log 2>&1; tail -2 gpurun_out/r2g_fwd.log
