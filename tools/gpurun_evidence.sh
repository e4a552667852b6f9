# Developer tool: the command handed to gpurun for an evidence run; edited per run, outputs under gpurun_out/ (the summaries
# worth keeping are copied to profiles/ by hand).  This version: per-config fwd+bwd table, ours vs the unmodified reference.
mkdir -p gpurun_out
timeout 300 python tools/config_table.py --out gpurun_out/r1z_config_table.json > gpurun_out/r1z_config_table.log 2>&1; tail -8 gpurun_out/r1z_config_table.log
