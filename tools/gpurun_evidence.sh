mkdir -p gpurun_out
T=r2bk
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${T}_bench_ours_2gpu.json 2> gpurun_out/${T}_bench_ours_2gpu.err; tail -c 700 gpurun_out/${T}_bench_ours_2gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/${T}_bench_reference_2gpu.json 2> gpurun_out/${T}_bench_reference_2gpu.err; tail -c 400 gpurun_out/${T}_bench_reference_2gpu.json
