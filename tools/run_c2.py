"""Development aid: run a few sweeps of the bench workload (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
args = sys.argv[1:]
sys.argv = [sys.argv[0]]
import bench
import dpmm_pkg
pkg = dpmm_pkg.load()
case = bench.build_case(args[0] if args else "c2")
g = pkg.GpuSweep(case["x"], case["kind"], seed=1)
bench.set_params(g, case)
for _ in range(int(args[1]) if len(args) > 1 else 4):
    g.sample_labels(False); g.sample_sublabels(); g.suff_stats(fetch=False)
g.sync()
g.close()
