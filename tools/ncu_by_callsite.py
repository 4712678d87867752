"""Development aid: like ncu_by_line.py, but every SASS instruction is charged to the OUTERMOST source line of
the chosen file in its inlining chain (nvdisasm -gi), so that helper calls (mbarrier waits, Philox, tcgen05
wrappers) show up at the line of the kernel that made them.
usage: ncu_by_callsite.py <source.csv (one kernel section)> <nvdisasm -gi output> <mangled function> <file name>"""
import csv, re, sys, collections
src_csv, dis, fn, own = sys.argv[1:5]
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + fn))
cur = None
pending = []
off2line = {}
for l in lines[start + 1:]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        pending.append(m.groups())
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", l)
    if m:
        if pending:
            # the chain: innermost first; take the last entry that names the own file (either side)
            key = None
            for f, ln, f2, ln2 in pending:
                if f.endswith(own):
                    key = int(ln)
                if f2 and f2.endswith(own):
                    key = int(ln2)
            if key is not None:
                cur = key
            pending = []
        off2line[int(m.group(1), 16)] = cur
    if l.startswith(".text.") and off2line:
        break
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for r in rows[2:]:
    if len(r) <= iinst or r[ia] == "Address":
        continue
    off = int(r[ia], 16) - base
    a = agg[off2line.get(off)]
    a[0] += int(r[isamp] or 0)
    a[1] += int(r[iinst] or 0)
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            a[2][hdr[i]] += v
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
print(f"total samples {tot_s}, instructions {tot_i}")
for key, a in sorted(agg.items(), key=lambda kv: (kv[0] or 0)):
    if a[0] * 300 < tot_s and a[1] * 300 < tot_i:
        continue
    top = ", ".join(f"{k[6:]}={v}" for k, v in a[2].most_common(3))
    print(f"{own}:{key}  samples {a[0]:6d} ({100*a[0]/tot_s:4.1f}%)  inst {a[1]:9d} ({100*a[1]/tot_i:4.1f}%)  {top}")
