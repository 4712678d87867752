"""Development aid: repeated full-size fits (hang / divergence soak).  usage: soak_fit.py <seconds> [D]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dpmm_pkg
pkg = dpmm_pkg.load()
from dpmmsubclusters_jl_b200.host import normalized_mutual_info
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
D = int(sys.argv[2]) if len(sys.argv) > 2 else 32
n, K = (1_000_000, 20) if D == 32 else (400_000, 12)
t0 = time.time(); runs = 0
while time.time() - t0 < budget:
    seed = runs + 1
    x, z, _, _ = pkg.generate_gaussian_data(n, D, K, 100.0 if runs % 3 else 4.0, np.random.default_rng(seed))
    out = pkg.fit(x, 10.0, iters=100, seed=seed, burnout=20 if runs % 2 else 8, smart_splits=(runs % 4 == 3))
    runs += 1
    print(f"run {runs}: D={D} K={len(out[1])} NMI={normalized_mutual_info(z, out[0]):.3f} loop {sum(out[3]):.3f}s", flush=True)
print("soak done:", runs, "fits")
