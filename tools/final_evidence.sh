mkdir -p gpurun_out
T=r2n
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${T}_gpu_tests.log; cat gpurun_out/${T}_gpu_tests.log
timeout 400 python bench.py --steps 200 --warmup 10 2> gpurun_out/${T}_bench_c2.err | grep "^{" > gpurun_out/${T}_bench_c2.json
timeout 300 python bench.py --workload c5s --steps 50 --warmup 5 --no-cpu-baseline --no-fit 2>/dev/null | grep "^{" > gpurun_out/${T}_bench_c5s.json
timeout 300 python bench.py --workload c2 --state overlap --steps 50 --warmup 5 --no-cpu-baseline --no-fit 2>/dev/null | grep "^{" > gpurun_out/${T}_bench_c2_overlap.json
PROBE_GT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gauss_label_tc2|niw_substats_tc" -s 8 -c 2 -o gpurun_out/${T}_prof python tools/tc2_probe.py c2 3 > gpurun_out/${T}_ncu.log 2>&1
PROBE_GT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gauss_label_tc2|niw_sublabel_tc64|niw_stats_tc64" -s 9 -c 3 -o gpurun_out/${T}_prof_c5s python tools/tc2_probe.py c5s 2 > gpurun_out/${T}_ncu_c5s.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/${T}_launches_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fit > /dev/null 2>&1
ls -la gpurun_out/${T}_*
