# round-1 final evidence run (gpurun): both bench arms back to back, the two-pass tool, k_recolor under ncu
mkdir -p gpurun_out
timeout 200 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r1t_bench_reference.json 2> gpurun_out/r1t_bench_reference.err; tail -c 300 gpurun_out/r1t_bench_reference.json
timeout 200 python bench.py > gpurun_out/r1t_bench_ours.json 2> gpurun_out/r1t_bench_ours.err; head -c 300 gpurun_out/r1t_bench_ours.json; echo
timeout 150 python tools/two_pass_times.py --out gpurun_out/r1t_two_pass.json > gpurun_out/r1t_two_pass.log 2>&1; tail -2 gpurun_out/r1t_two_pass.log
timeout 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_recolor -c 8 --csv --log-file gpurun_out/r1t_recolor.csv python tools/two_pass_times.py --views 2 > gpurun_out/r1t_recolor.log 2>&1; tail -12 gpurun_out/r1t_recolor.csv
