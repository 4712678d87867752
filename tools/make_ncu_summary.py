"""Development aid: profiles/<tag>_ncu_summary.md from gpurun_out/<tag>_prof.ncu-rep, the launch list and the bench line.
usage: make_ncu_summary.py <tag> "<note>" """
import csv, subprocess, collections, json, sys
tag, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = list(csv.reader(open(f'profiles/{tag}_launches_c2.csv')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
d = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > 5:
        try: d[r[4].split('(')[0]].append(float(r[-1].replace(',', '')))
        except ValueError: pass
tot = sum(sum(v) for v in d.values())
table = "\n".join(f"| `{k[:60]}` | {len(v)} | {sum(v)/len(v)/1000:.1f} | {sum(v)/tot:.3f} |" for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])))
out = subprocess.run(f"ncu -i gpurun_out/{tag}_prof.ncu-rep --page raw --csv", shell=True, capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines())); h = r[0]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']
b = json.load(open(f'profiles/{tag}_bench_c2.json'))
md = [f"# ncu --set full summary, round 1 (capture {tag[-1]}): tools/run_c2.py c2 6, C2 (N=1e6, D=32, K=20), one launch per kernel", "",
      f"Raw report: gpurun_out/{tag}_prof.ncu-rep (scratch, not committed).  Launch list of `bench.py --steps 5 --warmup 3`: {tag}_launches_c2.csv; bench line of the same build: {tag}_bench_c2.json.",
      note, ""]
traffic = {}
for row in r[2:]:
    name = row[h.index('Kernel Name')]
    md += [f"## {name}", "", "| metric | value | unit |", "|---|---|---|"]
    for k in keys:
        if k in h:
            i = h.index(k); md.append(f"| {k} | {row[i]} | {r[1][i]} |")
    md.append("")
    f = lambda v, u: v * {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1}[u]
    rd = f(float(row[h.index('dram__bytes_read.sum')]), r[1][h.index('dram__bytes_read.sum')])
    wr = f(float(row[h.index('dram__bytes_write.sum')]), r[1][h.index('dram__bytes_write.sum')])
    traffic['label' if 'label' in name else 'sublabel_stats_fused'] = rd + wr
st = b['stages']
md += ["## Launch list (ncu --metrics gpu__time_duration.sum, serialised, cold cache): share of the step", "",
       "| kernel | launches | avg us | share |", "|---|---|---|---|", table, "",
       f"`niw_pack_kernel` belongs to `dpmm_set_params_niw` (the e2e arm and the set-up), not to the device-resident step.  Inside the step the label kernel and the fused sub-label + statistics kernel are the two dominant launches here and by the CUDA-event timers of {tag}_bench_c2.json ({st['label']['us_per_step']:.1f} / {st['sublabel']['us_per_step']:.1f} of {b['ms_per_step']*1e3:.0f} us): the shares agree.", ""]
open(f'profiles/{tag}_ncu_summary.md', 'w').write("\n".join(md))
t = json.load(open('profiles/traffic.json'))
t['c2'].update(traffic)
t['source'] = f"profiles/{tag}_ncu_summary.md (label, fused sub-label+statistics: dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full); sublabel/stats = the separate kernels, profiles/r1b_ncu_summary.md"
json.dump(t, open('profiles/traffic.json', 'w'), indent=1)
print(table); print(traffic)
