mkdir -p gpurun_out
T=r2bn
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "accumulate or two_streams or fusion" > gpurun_out/${T}_pytest.txt 2>&1; tail -5 gpurun_out/${T}_pytest.txt | cut -c1-200
timeout 600 python bench.py --quick --per-stream-buffers > gpurun_out/${T}_bench_perstream.json 2> gpurun_out/${T}_bench_perstream.err; python -c "
import json; d=json.loads(open('gpurun_out/${T}_bench_perstream.json').read().strip().splitlines()[-1]); print('per-stream', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['stages']['preprocess_bwd'])"
timeout 600 python bench.py --quick > gpurun_out/${T}_bench_onebuf.json 2> gpurun_out/${T}_bench_onebuf.err; python -c "
import json; d=json.loads(open('gpurun_out/${T}_bench_onebuf.json').read().strip().splitlines()[-1]); print('one-buffer', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['stages']['preprocess_bwd'], d['roofline'].get('traffic'), d['roofline'].get('issue_frac'))"
tail -3 gpurun_out/${T}_bench_onebuf.err
