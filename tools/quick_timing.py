"""Quick per-kernel timing of one sweep at a given shape (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dpmm_pkg
from tests.util import make_niw_case, make_mnm_case, set_params

pkg = dpmm_pkg.load()
kind = sys.argv[1] if len(sys.argv) > 1 else "niw"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
D = int(sys.argv[3]) if len(sys.argv) > 3 else 32
K = int(sys.argv[4]) if len(sys.argv) > 4 else 20
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 10
t0 = time.time()
case = make_niw_case(D, K, n, 1) if kind == "niw" else make_mnm_case(D, K, n, 1)
print(f"case built in {time.time()-t0:.1f}s", flush=True)
g = pkg.GpuSweep(case["x"], case["kind"], seed=1)
set_params(g, case)
for _ in range(3):
    g.sample_labels(); g.sample_sublabels(); g.suff_stats(fetch=False)
g.sync()
g.timing_enable(True)
t0 = time.time()
for _ in range(iters):
    g.sample_labels(); g.sample_sublabels(); g.suff_stats(fetch=False)
g.sync()
wall = (time.time() - t0) / iters * 1e3
t = g.timing_read()
print(f"{kind} n={n} D={D} K={K}: wall {wall:.3f} ms/iter (timers on)")
for k, (ms, cnt) in t.items():
    if cnt:
        print(f"  {k:10s} {ms/iters*1e3:10.1f} us/iter  ({cnt//iters} launches/iter)")
g.timing_enable(False)
t0 = time.time()
for _ in range(iters):
    g.sample_labels(); g.sample_sublabels(); g.suff_stats(fetch=False)
g.sync()
print(f"  wall without timers: {(time.time()-t0)/iters*1e3:.3f} ms/iter")
g.close()
