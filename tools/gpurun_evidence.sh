mkdir -p gpurun_out
T=r2bq
timeout 600 python tools/stage_times.py 2>&1 | grep "tile_sort\|sum of" 
timeout 600 python tools/stage_times.py --P 200000 2>&1 | grep "tile_sort\|sum of"
timeout 600 python tools/stage_times.py --P 4000000 --W 3840 --H 2160 2>&1 | grep "tile_sort\|sum of"
timeout 600 python tools/stage_times.py --random 2>&1 | grep "tile_sort\|sum of"
timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -q -x > gpurun_out/${T}_pytest.txt 2>&1; tail -3 gpurun_out/${T}_pytest.txt | cut -c1-200
