"""Developer tool: summarise ncu CSV exports.
    python tools/ncu_summary.py launches gpurun_out/x_launches.csv [views]
    python tools/ncu_summary.py full gpurun_out/x_full.ncu-rep        (needs ncu on PATH)
"""
import collections, csv, subprocess, sys, io

def launches(path, views=1):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H = rows[hdr]; rows = rows[hdr + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    d = collections.defaultdict(list)
    for r in rows:
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] in ("ns", "nsecond") else v * 1000 if r[ui] in ("ms", "msecond") else v
        d[r[ki][:60]].append(v)
    tot = sum(sum(v) for v in d.values())
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:60s} n={len(v):3d} avg={sum(v)/len(v):8.1f}us last={v[-1]:8.1f} share={100*sum(v)/tot:5.1f}%")
    print(f"total {tot:.1f} us  per view {tot/views:.1f} us")

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
STALL = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "barrier", "lg_throttle", "mio_throttle", "math_pipe_throttle", "not_selected",
          "no_instruction", "branch_resolving", "sleeping", "membar", "dispatch_stall", "drain", "imc_miss", "tex_throttle"]

def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H, U = rows[0], rows[1]
    for r in rows[2:]:
        print("-----")
        for w in WANT:
            if w in H:
                print(f"  {w}: {r[H.index(w)]} {U[H.index(w)]}")
        st = []
        for s in STALLS:
            n = STALL % s
            if n in H:
                st.append(f"{s}={float(r[H.index(n)]):.2f}")
        print("  stalls/issue: " + " ".join(st))

def source(path, kernel, top=40):
    """Per CUDA source line: instructions executed and stall samples (needs -lineinfo and --import-source on)."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
    H = rows[hi]
    ii, ti = H.index("Instructions Executed"), H.index("# Samples")
    agg = {}
    for r in rows[hi + 1:]:
        if len(r) > ii and r[0].strip().isdigit():
            try:
                k = (int(r[0]), r[1].strip())
                v = agg.setdefault(k, [0, 0])
                v[0] += int(r[ii]); v[1] += int(r[ti])
            except ValueError:
                pass
    tot_i = sum(v[0] for v in agg.values()) or 1
    tot_s = sum(v[1] for v in agg.values()) or 1
    print(f"total inst {tot_i}  samples {tot_s}")
    for (ln, src), (ins, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100*ins/tot_i:5.1f}% inst {100*smp/tot_s:5.1f}% smp  L{ln:>4d}  {src[:120]}")


if __name__ == "__main__" and sys.argv[1] != "regions":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 1)
    elif sys.argv[1] == "source":
        source(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
    else:
        full(sys.argv[2])


def regions(path, kernel, srcfile, markers):
    """Instructions and stall samples per source region. markers: substrings of source lines that START a region (in file order)."""
    import re
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
    H = rows[hi]
    idx = {h: i for i, h in enumerate(H)}
    stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
    src = open(srcfile).read().split("\n")
    starts = []
    for m in markers:
        ln = [i + 1 for i, l in enumerate(src) if m in l]
        starts.append((ln[0] if ln else 10**9, m))
    starts.sort()
    agg = [collections.Counter() for _ in starts]
    inst = [0] * len(starts)
    tinst = [0] * len(starts)
    for r in rows[hi + 1:]:
        if len(r) < len(H) or not r[0].strip().isdigit():
            continue
        ln = int(r[0])
        k = max([i for i, (s, _) in enumerate(starts) if s <= ln], default=None)
        if k is None:
            continue
        try:
            inst[k] += int(r[idx["Instructions Executed"]]); tinst[k] += int(r[idx["Thread Instructions Executed"]])
        except ValueError:
            pass
        for s in stalls:
            try:
                agg[k][s] += int(r[idx[s]])
            except ValueError:
                pass
    ti = sum(inst) or 1
    ts = sum(sum(a.values()) for a in agg) or 1
    for (s, m), n, tn, a in zip(starts, inst, tinst, agg):
        tot = sum(a.values())
        top = " ".join(f"{k.replace('stall_', '')}={100 * v / max(tot, 1):.0f}%" for k, v in a.most_common(5))
        print(f"L{s:4d} {m[:38]:38s} inst {100 * n / ti:5.1f}% ({n / 1e6:6.2f}M, {tn / max(n, 1):4.1f} lanes)  samples {100 * tot / ts:5.1f}%  [{top}]")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "regions":
    regions(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5:])
