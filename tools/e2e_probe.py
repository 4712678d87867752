import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench, dpmm_pkg
pkg = dpmm_pkg.load()
case = bench.build_case("c2", 0, 0)
g = pkg.GpuSweep(case["x"], case["kind"], seed=1)
bench.set_params(g, case)
for _ in range(5):
    g.sample_labels(False); g.sample_sublabels(); out = g.suff_stats()
def t(fn, n=100):
    g.sync(); t0 = time.perf_counter()
    for _ in range(n): fn()
    g.sync(); return (time.perf_counter() - t0) / n * 1e6
res = {}
res["set_params+sync"] = t(lambda: (bench.set_params(g, case), g.sync()))
res["set_params (async)"] = t(lambda: bench.set_params(g, case))
res["sweep nofetch+sync"] = t(lambda: (g.sample_labels(False), g.sample_sublabels(), g.suff_stats(fetch=False), g.sync()))
res["sweep fetch"] = t(lambda: (g.sample_labels(False), g.sample_sublabels(), g.suff_stats(out=out)))
res["e2e_hp"] = t(lambda: (bench.set_params(g, case), g.sample_labels(False), g.sample_sublabels(), g.suff_stats(out=out)))
g.timing_enable(True)
for _ in range(50): bench.set_params(g, case)
tim = g.timing_read(); res["params_kernels_us"] = tim["params"][0] / 50 * 1e3
print(json.dumps({k: round(v, 1) for k, v in res.items()}))
