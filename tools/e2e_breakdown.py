"""Development aid: where does the end-to-end (host-buffer) step time go?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0]]
import numpy as np
import bench, dpmm_pkg
pkg = dpmm_pkg.load()
case = bench.build_case("c2")
g = pkg.GpuSweep(case["x"], case["kind"], seed=1)
def t(fn, n=50):
    for _ in range(5): fn()
    g.sync(); t0 = time.perf_counter()
    for _ in range(n): fn()
    g.sync(); return (time.perf_counter() - t0) / n * 1e6
bench.set_params(g, case)
print("set_params            %.1f us" % t(lambda: bench.set_params(g, case)))
print("sample_labels         %.1f us" % t(lambda: g.sample_labels(False)))
print("sample_sublabels      %.1f us" % t(lambda: g.sample_sublabels()))
print("suff_stats(nofetch)   %.1f us" % t(lambda: g.suff_stats(fetch=False)))
print("suff_stats(fetch)     %.1f us" % t(lambda: g.suff_stats()))
def step():
    bench.set_params(g, case); g.sample_labels(False); g.sample_sublabels(); g.suff_stats()
print("full e2e step         %.1f us" % t(step))
import ctypes as C
mu = np.ascontiguousarray(case["mu"]); inv = np.ascontiguousarray(case["inv_sigma"]); ld = np.ascontiguousarray(case["logdet"]); w = case["weights"]; lr = np.ascontiguousarray(case["lr_weights"])
f = C.c_float
from dpmmsubclusters_jl_b200.sweep import _ptr
raw = lambda: g.lib.dpmm_set_params_niw(g.h, case["K"], _ptr(mu, f), _ptr(inv, f), _ptr(ld, f), _ptr(w, f), _ptr(lr, f))
print("set_params raw C call %.1f us" % t(raw))
