import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench, dpmm_pkg
pkg = dpmm_pkg.load()
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
case = bench.build_case(name, 0, 0)
K = case["K"]
g = pkg.GpuSweep(case["x"], case["kind"], seed=1)
bench.set_params(g, case)
hy = case["hyper"]; g.set_hyper_niw(hy.κ, hy.m, hy.ν, hy.ψ, 10.0)
sp = np.ones(K, bool)
for _ in range(3):
    g.sample_labels(False); g.sample_sublabels()
g.posterior_step(None)
def it():
    g.sample_params(K); g.sample_labels(False); g.sample_sublabels(); return g.posterior_step(None, splittable=sp)
for _ in range(5): it()
def t(fn, n=100):
    g.sync(); t0 = time.perf_counter()
    for _ in range(n): fn()
    g.sync(); return (time.perf_counter() - t0) / n * 1e6
res = {"iteration": t(it), "sample_params+sync": t(lambda: (g.sample_params(K), g.sync())),
       "posterior_step(from sweep)": None}
res["sweep+post"] = t(lambda: (g.sample_labels(False), g.sample_sublabels(), g.posterior_step(None, splittable=sp)))
res["sweep+post no merge"] = t(lambda: (g.sample_labels(False), g.sample_sublabels(), g.posterior_step(None)))
res["sweep nofetch+sync"] = t(lambda: (g.sample_labels(False), g.sample_sublabels(), g.suff_stats(fetch=False), g.sync()))
print(json.dumps({k: (round(v, 1) if v else v) for k, v in res.items()}))
