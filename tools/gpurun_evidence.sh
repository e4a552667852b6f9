mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_refine.py > gpurun_out/r2aw_sharded_refine.log 2>&1; echo "rc=$?" >> gpurun_out/r2aw_sharded_refine.log; tail -25 gpurun_out/r2aw_sharded_refine.log | cut -c1-400
