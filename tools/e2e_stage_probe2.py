"""Development aid: which part of the complete iteration slows the D = 64 label kernel at 12.5e6 points."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, dpmm_pkg
pkg = dpmm_pkg.load()
case = bench.build_case(sys.argv[1] if len(sys.argv) > 1 else "c5", 0, 0)
g = pkg.GpuSweep(case["x"], case["kind"], seed=1)
bench.set_params(g, case)
g.set_labels(case["gt"])
hy = case["hyper"]
g.set_hyper_niw(hy.κ, hy.m, hy.ν, hy.ψ, case["alpha"])
g.sample_labels(False); g.sample_sublabels(); g.posterior_step(None)
sp = np.ones(case["K"], bool)
def run(tag, fn, n=6):
    for _ in range(10): fn()
    g.sync(); g.timing_enable(True)
    for _ in range(n): fn()
    t = g.timing_read(); g.timing_enable(False)
    print(tag, {k: round(v[0] / n * 1e3, 1) for k, v in t.items() if v[1]}, g.tc_stats(overflow=True), flush=True)
def frozen(): g.sample_labels(False); g.sample_sublabels(); g.suff_stats(fetch=False)
def frozen_post(): g.sample_labels(False); g.sample_sublabels(); g.posterior_step(None, splittable=sp)
def sampled(): g.sample_params(case["K"]); g.sample_labels(False); g.sample_sublabels(); g.posterior_step(None, splittable=sp)
run("frozen sweep      ", frozen)
run("frozen + posterior", frozen_post)
run("sampled params    ", sampled)
mu, lf, ld, w, lr = g.get_params_niw(case["K"])
print("weights: sampled min/max", float(w.min()), float(w.max()), " frozen min/max", float(case["weights"].min()), float(case["weights"].max()))
print("logdet cluster: sampled", ld[:3, 0], " frozen", case["logdet"][:3, 0])
bench.set_params(g, case)
run("frozen again      ", frozen)
