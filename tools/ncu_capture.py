"""Developer tool: turn an `ncu --set full` capture of the headline workload into profiles/ncu_capture.json -- per kernel the DRAM
traffic and warp instructions per launch that bench.py prints as roofline.traffic / issue_frac -- stamped with the sha of the kernel
sources the capture was taken from (bench.py withholds the numbers when the sources have changed since).

    python tools/ncu_capture.py gpurun_out/r2_full.ncu-rep [name-of-capture]
"""
import collections, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util
spec = importlib.util.spec_from_file_location("_bench", os.path.join(ROOT, "bench.py"))

def sources_sha():
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "gaustar_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if not os.path.isfile(os.path.join(d, f)) or f.startswith("."):
            continue
        h.update(f.encode()); h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]

path = sys.argv[1]
out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
H, U = rows[0], rows[1]
col = {h: i for i, h in enumerate(H)}
def val(r, name):
    v = float(r[col[name]].replace(",", "")); u = U[col[name]]
    return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "inst": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u, 1.0)
agg = collections.defaultdict(list)
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0]
    agg[name].append((val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"), val(r, "smsp__inst_executed.sum"), val(r, "gpu__time_duration.sum")))
kernels = {k: {"dram_bytes": int(sum(x[0] for x in v) / len(v)), "warp_inst": int(sum(x[1] for x in v) / len(v)),
               "isolated_us": round(sum(x[2] for x in v) / len(v), 1), "launches": len(v)} for k, v in agg.items()}
doc = {"capture": sys.argv[2] if len(sys.argv) > 2 else os.path.basename(path), "sources_sha": sources_sha(), "kernels": kernels}
json.dump(doc, open(os.path.join(ROOT, "profiles", "ncu_capture.json"), "w"), indent=1)
print(json.dumps(doc, indent=1))
