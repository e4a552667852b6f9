# The round-2 evidence run (profiles/r2bp_*): `gpurun --timeout 4000 -- 'bash tools/gpurun_evidence.sh'`, then on the build host
#   python tools/ncu_capture.py gpurun_out/r2bp_full.ncu-rep r2bp_full      -> profiles/ncu_capture.json (stamped with the sources' sha)
#   python tools/ncu_summary.py launches|full ...                           -> profiles/r2bp_launches_summary.txt, r2bp_ncu_full_summary.txt
# Other measurements of the round: tools/stage_times.py [--dual], tools/two_pass_times.py [--passes 3], tools/config_table.py,
# tools/gaustar_iteration_times.py, tools/elementwise_noise.py, tools/sharded_refine.py (2 GPUs), tools/allreduce_probe.py (8 GPUs).
mkdir -p gpurun_out
T=r2bp
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.txt; tail -4 gpurun_out/${T}_pytest.txt | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.txt 2>&1; tail -2 gpurun_out/${T}_smoke.txt | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -c 40 -o gpurun_out/${T}_full python tools/profile_view.py --views 3 > gpurun_out/${T}_full.log 2>&1; tail -1 gpurun_out/${T}_full.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench_steps1.csv python bench.py --steps 1 --warmup 3 --quick > gpurun_out/${T}_launches.log 2>&1; tail -1 gpurun_out/${T}_launches.log | cut -c1-200
timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -c 300 gpurun_out/${T}_bench_reference.json
timeout 900 python bench.py > gpurun_out/${T}_bench_ours.json 2> gpurun_out/${T}_bench_ours.err; tail -c 1200 gpurun_out/${T}_bench_ours.json
