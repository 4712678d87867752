"""Development aid: per-kind device times of the complete iteration (parameter step + sweep + posterior step)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, dpmm_pkg
pkg = dpmm_pkg.load()
name = sys.argv[1] if len(sys.argv) > 1 else "c5s"
case = bench.build_case(name, 0, 0)
g = pkg.GpuSweep(case["x"], case["kind"], seed=1)
bench.set_params(g, case)
g.set_labels(case["gt"])
hy = case["hyper"]
g.set_hyper_niw(hy.κ, hy.m, hy.ν, hy.ψ, case["alpha"])
g.sample_labels(False); g.sample_sublabels()
g.posterior_step(None)
sp = np.ones(case["K"], bool)
def it():
    g.sample_params(case["K"]); g.sample_labels(False); g.sample_sublabels()
    return g.posterior_step(None, splittable=sp)
for _ in range(12):
    it()
g.sync()
t0 = time.perf_counter()
for _ in range(10):
    it()
wall = (time.perf_counter() - t0) / 10 * 1e3
g.timing_enable(True)
for _ in range(10):
    it()
t = g.timing_read()
print(name, "wall ms/iter", round(wall, 3), {k: round(v[0] / 10 * 1e3, 1) for k, v in t.items() if v[1]}, "tc_stats", g.tc_stats(overflow=True))
