mkdir -p gpurun_out
timeout 900 python tools/gaustar_iteration_times.py > gpurun_out/r2aq_iteration_times.json 2> gpurun_out/r2aq_iteration_times.err; tail -3 gpurun_out/r2aq_iteration_times.err; cat gpurun_out/r2aq_iteration_times.json
