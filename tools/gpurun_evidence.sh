mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "not full_size" > gpurun_out/r2ao_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ao_pytest.log; tail -3 gpurun_out/r2ao_pytest.log
for args in "" "--P 200000" "--P 4000000 --W 3840 --H 2160 --views 3"; do echo "== new $args"; timeout 300 python tools/stage_times.py $args 2>&1 | grep -E "blend_fwd|rror"; done
GSTAR_LIB_PATH=gaustar_b200/lib/variants/dbg.so timeout 300 python tools/stage_times.py --views 1 2>&1 | grep -E "blk (0|600) warp (0|7) cyc|blk (0|20|600) n=" | tail -8
