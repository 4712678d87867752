#!/bin/bash
# round-2 GPU pass A: new label kernel parity + first timings
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "tc2" 2>&1 | tail -40 > gpurun_out/a_tc2.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "full_sweep or refine or argmax or full_size or fused" 2>&1 | tail -40 > gpurun_out/a_parity.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-fit > gpurun_out/a_bench_c2.json 2> gpurun_out/a_bench_c2.err
DPMM_LABEL_TC=1 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-fit > gpurun_out/a_bench_c2_old.json 2> gpurun_out/a_bench_c2_old.err
timeout 300 python bench.py --workload c5s --steps 20 --warmup 3 --no-cpu-baseline --no-fit > gpurun_out/a_bench_c5s.json 2> gpurun_out/a_bench_c5s.err
tail -5 gpurun_out/a_tc2.log gpurun_out/a_parity.log
cat gpurun_out/a_bench_c2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stages'], d['roofline'])"
