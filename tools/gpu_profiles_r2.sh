#!/bin/bash
# round-2 evidence pass: bench lines of every workload, launch list, full ncu captures of the dominant kernels
mkdir -p gpurun_out
T=${1:-r2k}
timeout 400 python bench.py --steps 200 --warmup 10 > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2>/dev/null
timeout 300 python bench.py --workload c2 --state overlap --steps 50 --warmup 5 --no-cpu-baseline --no-fit > gpurun_out/${T}_bench_c2_overlap.json 2>/dev/null
for w in c1 c3 c4 c5s; do
  timeout 600 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline --no-fit > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/${T}_launches_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fit > /dev/null 2>&1
PROBE_GT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gauss_label_tc2|niw_substats_tc" -s 8 -c 2 -o gpurun_out/${T}_prof python tools/tc2_probe.py c2 3 > gpurun_out/${T}_ncu.log 2>&1
PROBE_GT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gauss_label_tc2|niw_sublabel_tc64|niw_stats_tc64" -s 9 -c 3 -o gpurun_out/${T}_prof_c5s python tools/tc2_probe.py c5s 2 > gpurun_out/${T}_ncu_c5s.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"mnm_label_tc|gauss_label_warp" -s 3 -c 1 -o gpurun_out/${T}_prof_c3 python tools/tc2_probe.py c3 2 > gpurun_out/${T}_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"gauss_label_warp|gauss_label_kernel" -s 3 -c 1 -o gpurun_out/${T}_prof_c4 python tools/tc2_probe.py c4 2 > gpurun_out/${T}_ncu_c4.log 2>&1
ls -la gpurun_out/${T}_*
