"""fit() on a bench case: iterations/s, final K, NMI (device vs host parameter step)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, dpmm_pkg
pkg = dpmm_pkg.load()
from dpmmsubclusters_jl_b200.host import normalized_mutual_info
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
mode = (sys.argv[2] != "host") if len(sys.argv) > 2 else True
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 100
case = bench.build_case(name, 0, 0)
for rep in range(2):
    t0 = time.perf_counter()
    out = pkg.fit(case["x"], 10.0, iters=iters, seed=1 + rep, burnout=20, device_params=mode)
    dt = time.perf_counter() - t0
    print(json.dumps({"case": name, "device_params": mode, "iters": iters, "seconds": dt, "iters_per_s": iters / dt,
                      "sum_iter_s": sum(out[3]), "final_K": len(out[1]), "nmi": normalized_mutual_info(case["gt"], out[0]),
                      "k_hist": out[6][::10]}))
