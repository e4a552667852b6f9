# Developer tool: the command handed to gpurun for an evidence run; edited per run, outputs under gpurun_out/ (the summaries
# worth keeping are copied to profiles/ by hand).  This version: per-stage times at config #2's shape.
mkdir -p gpurun_out
timeout 120 python tools/stage_times.py --P 200000 --views 6 > gpurun_out/r1z_stage_200k.log 2>&1; tail -12 gpurun_out/r1z_stage_200k.log
