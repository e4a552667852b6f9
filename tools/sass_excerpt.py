"""Developer tool: per-kernel counts of the SASS mnemonics that show what the kernels are made of (TMA bulk copies, mbarriers, vector
reductions, shared-memory atomics, warp reductions) -> profiles/<name>.  Run after a build:  python tools/sass_excerpt.py r2_sass_excerpt.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "gaustar_b200", "lib", "libgstar_raster.so")], capture_output=True, text=True).stdout
cur, hits = None, collections.OrderedDict()
pat = re.compile(r"\b(UBLKCP|SYNCS|REDG|RED\.|ATOMS|ATOMG|UTMALDG|LDGSTS|REDUX|CREDUX|MATCH|BAR\.|MUFU\.EX2|LDG\.E\.128\.CONSTANT|STG\.E\.128)[\w\.]*")
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); hits[cur] = collections.Counter(); continue
    if cur:
        m = pat.search(line)
        if m:
            hits[cur][m.group(0)] += 1
dst = os.path.join(ROOT, "profiles", sys.argv[1] if len(sys.argv) > 1 else "sass_excerpt.txt")
with open(dst, "w") as f:
    f.write("SASS mnemonics per kernel of gaustar_b200/lib/libgstar_raster.so (cuobjdump -sass; counts of static instructions; tools/sass_excerpt.py).\n"
            "UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier ops, REDG...F32x4 = red.global.add.v4.f32, ATOMS = shared-memory atomics,\n"
            "REDUX/CREDUX = warp reductions, MATCH = __match_any_sync.\n\n")
    for k, c in hits.items():
        if not k.startswith("_ZN5gstar"):
            continue
        name = re.sub(r"ENS_.*$|E[PKi].*$", "", re.sub(r"^_ZN5gstar\d+", "", k))
        f.write(f"{name}: " + ", ".join(f"{m} x{n}" for m, n in sorted(c.items())) + "\n")
print(open(dst).read())
