mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2al_bench_ours.json 2> gpurun_out/r2al_bench_ours.err ) 2>&1 | grep real
tail -3 gpurun_out/r2al_bench_ours.err
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2al_bench_reference.json 2> gpurun_out/r2al_bench_reference.err ) 2>&1 | grep real
tail -3 gpurun_out/r2al_bench_reference.err
python - <<'PY'
import json
for n in ("ours","reference"):
    try:
        d=json.load(open(f"gpurun_out/r2al_bench_{n}.json"))
        print(n, "value", d["value"], "e2e", d["e2e"], "\n  refine", d.get("refine_step"), "\n  cpu", d.get("cpu_baseline"), d.get("cpu_baseline_naive"))
        if "roofline" in d: print("  roofline", {k:v for k,v in d["roofline"].items() if k!="stages"}); print("  stages", {k:(v["ms"],v["frac"],v.get("issue_frac")) for k,v in d["roofline"]["stages"].items()})
    except Exception as e: print(n, "ERR", e)
PY
