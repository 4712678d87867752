import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench, dpmm_pkg
pkg = dpmm_pkg.load()
case = bench.build_case("c2", 0, 0)
pkg.fit(case["x"], 10.0, iters=100, seed=1, burnout=20)
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter(); out = pkg.fit(case["x"], 10.0, iters=100, seed=2, burnout=20); print("fit s", time.perf_counter() - t0)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
