import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, dpmm_pkg
pkg = dpmm_pkg.load()
from dpmmsubclusters_jl_b200 import host as H
x, labels, _, _ = pkg.generate_gaussian_data(10 ** 5, 32, 20, 100.0, np.random.default_rng(5))
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    out = H.fit(x, 10.0, iters=60, seed=seed, burnout=10, device_params=True)
    print(seed, len(out[1]), H.normalized_mutual_info(labels, out[0]), flush=True)
