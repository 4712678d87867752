"""Development aid: end-to-end fit() on C2-shaped data (host parameter sampling in Python included)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dpmm_pkg
pkg = dpmm_pkg.load()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
x, z, _, _ = pkg.generate_gaussian_data(n, 32, 20, 100.0, np.random.default_rng(0))
t0 = time.perf_counter()
out = pkg.fit(x, 10.0, iters=iters, seed=1, gt=None, burnout=20)
dt = time.perf_counter() - t0
from dpmmsubclusters_jl_b200.host import normalized_mutual_info
print(f"fit: {iters} iterations in {dt:.2f} s ({iters/dt:.1f} iters/s incl. upload + Python host), K={len(out[1])}, "
      f"NMI={normalized_mutual_info(z, out[0]):.4f}, K history {out[6][::10]}")
print("per-iteration group_step times (ms), every 10th:", [round(t*1e3,1) for t in out[3][::10]])
