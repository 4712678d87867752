"""Development aid: randomised parity soak of the D=32 path (tcgen05 label kernel + fused sub-label/statistics
kernel) against the oracle: random K, n, spread, seeds; full compare_sweeps each time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dpmm_pkg
from oracle import dpmm_oracle as O
from tests.util import make_niw_case, compare_sweeps
pkg = dpmm_pkg.load()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
t0 = time.time(); runs = 0; redone_tot = 0
while time.time() - t0 < budget:
    K = int(rng.choice([1, 2, 3, 5, 8, 13, 20, 23, 31, 60]))
    n = int(rng.choice([130, 1000, 5000, 20000, 60000]))
    spread = float(rng.choice([0.0, 0.5, 2.5, 10.0, 40.0]))
    seed = int(rng.integers(1, 10**6))
    case = make_niw_case(32, K, n, seed=seed, spread=spread)
    g = pkg.GpuSweep(case["x"], pkg.NIW, seed=seed)
    o = O.OracleSweep(case["x"], O.NIW, seed=seed)
    try:
        rep = compare_sweeps(g, o, case, np.random.default_rng(seed), final=bool(rng.integers(0, 2)))
    except AssertionError as e:
        print(f"FAIL K={K} n={n} spread={spread} seed={seed}: {str(e)[:300]}", flush=True)
        raise
    f = g.fused_stats(); g.close()
    redone_tot += f[2]; runs += 1
    print(f"ok K={K:3d} n={n:6d} spread={spread:5.1f} seed={seed:7d} ties={rep['label_ties']},{rep['sub_ties']} stats_err={rep['stats_err']:.1e} fused={f}", flush=True)
print(f"{runs} configurations passed; exact recomputations {redone_tot}")
