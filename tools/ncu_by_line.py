"""Development aid: join an ncu SASS source page (CSV) with nvdisasm -g line info and aggregate
executed instructions and stall samples per source line.
usage: ncu_by_line.py <source.csv> <nvdisasm.txt> <function name> [file filter]"""
import csv, re, sys, collections
src_csv, dis, fn = sys.argv[1:4]
flt = sys.argv[4] if len(sys.argv) > 4 else None
# nvdisasm: sequence of (offset -> (file, line))
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + fn) or l.startswith(fn + ":"))
cur = ("?", 0)
off2line = {}
inl = []
for l in lines[start:]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
    if l.startswith("//-----") and off2line:
        break
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for r in rows[2:]:
    if len(r) <= iinst:
        continue
    off = int(r[ia], 16) - base
    key = off2line.get(off, ("?", 0))
    a = agg[key]
    a[0] += int(r[isamp] or 0)
    a[1] += int(r[iinst] or 0)
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            a[2][hdr[i]] += v
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
print(f"total samples {tot_s}, instructions {tot_i}")
for key, a in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if flt and flt not in key[0]:
        continue
    if a[0] * 200 < tot_s and a[1] * 200 < tot_i:
        continue
    top = ", ".join(f"{k[6:]}={v}" for k, v in a[2].most_common(3))
    print(f"{key[0]:28s}:{key[1]:4d}  samples {a[0]:6d} ({100*a[0]/tot_s:4.1f}%)  inst {a[1]:9d} ({100*a[1]/tot_i:4.1f}%)  {top}")
