"""profiles/r2_sass_evidence.md: count the tensor-core / TMA / TMEM mnemonics per kernel of the built library
(cuobjdump -sass).  usage: make_sass_evidence.py [out.md]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "dpmmsubclusters.jl_b200", "libdpmm_b200.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_evidence.md")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = ["UTCHMMA", "LDTM", "UTCBAR", "UTMALDG", "LDGSTS", "UTCATOMSWS", "FFMA2", "SYNCS", "HMMA"]
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        for n in names:
            if op == n or (n == "HMMA" and op.startswith("HMMA")):
                counts[cur][n] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
rows = sorted(zip(dem, counts.values()), key=lambda kv: (-kv[1]["UTCHMMA"], -kv[1]["LDGSTS"] - kv[1]["UTMALDG"], kv[0]))
md = ["# SASS evidence (cuobjdump -sass dpmmsubclusters.jl_b200/libdpmm_b200.so, round 2): tensor-core / TMA / TMEM mnemonics per kernel", "",
      "UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG = cp.async.bulk.tensor (TMA), LDGSTS = cp.async, "
      "UTCATOMSWS = tcgen05.alloc, FFMA2 = packed FP32 FMA, SYNCS = mbarrier ops; no HMMA (legacy mma.sync) anywhere.", "",
      "| kernel | " + " | ".join(names) + " |", "|---|" + "---|" * len(names)]
for name, c in rows:
    if sum(c.values()) == 0:
        continue
    md.append(f"| `{name[:100]}` | " + " | ".join(str(c[n]) for n in names) + " |")
open(out, "w").write("\n".join(md) + "\n")
print("\n".join(md[:16]))
