"""profiles/<tag>_ncu_summary.md from gpurun_out/<tag>_prof*.ncu-rep (ncu --set full captures) and the launch list.
usage: make_ncu_summary_r2.py <tag>"""
import csv, subprocess, collections, json, sys, os
tag = sys.argv[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']
md = [f"# ncu --set full summaries, round 2 (capture {tag})", "",
      "Raw reports: gpurun_out/*.ncu-rep (scratch, not committed).  `--clock-control none`; durations under the profiler are "
      "cold-cache and serialised -- the timed numbers are the CUDA-event ones of the bench lines in this directory.", ""]
traffic = {}
def unit_bytes(v, u):
    return float(v.replace(',', '')) * {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1, 'Tbyte': 1e12}[u]
for rep, case in [(f"gpurun_out/{tag}_prof.ncu-rep", "c2"), (f"gpurun_out/{tag}_prof_c5s.ncu-rep", "c5s"),
                  (f"gpurun_out/{tag}_prof_c3.ncu-rep", "c3"), (f"gpurun_out/{tag}_prof_c4.ncu-rep", "c4")]:
    if not os.path.exists(rep):
        continue
    out = subprocess.run(f"ncu -i {rep} --page raw --csv", shell=True, capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    if len(r) < 3:
        continue
    h = r[0]
    for row in r[2:]:
        name = row[h.index('Kernel Name')]
        md += [f"## {case}: `{name[:90]}`", "", "| metric | value | unit |", "|---|---|---|"]
        for k in keys:
            if k in h:
                i = h.index(k)
                md.append(f"| {k} | {row[i]} | {r[1][i]} |")
        rd = unit_bytes(row[h.index('dram__bytes_read.sum')], r[1][h.index('dram__bytes_read.sum')])
        wr = unit_bytes(row[h.index('dram__bytes_write.sum')], r[1][h.index('dram__bytes_write.sum')])
        md += [f"| DRAM traffic (read + write) | {(rd + wr) / 1e6:.1f} | MB |", ""]
        short = ("label_tc2" if "tc2" in name else "sublabel_stats_fused" if "substats" in name else "sublabel" if "sublabel" in name
                 else "stats" if "stats" in name else "label" if "label" in name else name[:20])
        traffic.setdefault(case, {})[short] = rd + wr
# launch list
p = f"gpurun_out/{tag}_launches_c2.csv"
if os.path.exists(p):
    rows = list(csv.reader(open(p)))
    hdr = [i for i, r_ in enumerate(rows) if r_ and r_[0] == 'ID'][0]
    d = collections.defaultdict(list)
    for r_ in rows[hdr + 1:]:
        if len(r_) > 5:
            try: d[r_[4].split('(')[0]].append(float(r_[-1].replace(',', '')))
            except ValueError: pass
    tot = sum(sum(v) for v in d.values())
    md += ["## Launch list of `bench.py --steps 5 --warmup 3` (ncu --metrics gpu__time_duration.sum, serialised, cold cache): share of the captured window", "",
           "| kernel | launches | avg us | share |", "|---|---|---|---|"]
    md += [f"| `{k[:70]}` | {len(v)} | {sum(v)/len(v)/1000:.1f} | {sum(v)/tot:.3f} |" for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1]))]
    md.append("")
open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(md))
t = json.load(open("profiles/traffic.json"))
for c, v in traffic.items():
    t.setdefault(c, {}).update(v)
t["source_r2"] = f"profiles/{tag}_ncu_summary.md: dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full"
json.dump(t, open("profiles/traffic.json", "w"), indent=1)
print("\n".join(md[:60]))
