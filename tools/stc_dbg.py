import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dpmm_pkg
from tests.util import make_niw_case, set_params
pkg = dpmm_pkg.load()
case = make_niw_case(32, 4, 20000, 1)
g = pkg.GpuSweep(case["x"], case["kind"], seed=1)
set_params(g, case)
g.sample_labels(); g.sample_sublabels()
os.environ["DPMM_STATS_TC"] = "0"
c0, sx0, sxx0 = g.suff_stats()
os.environ["DPMM_STATS_TC"] = "1"
os.environ["DPMM_STC_DEBUG"] = "1"
os.environ["DPMM_STC_MODE"] = sys.argv[1]
sys.stderr.write(f"=== mode {sys.argv[1]}\n")
c1, sx1, sxx1 = g.suff_stats()
lab = g.get_labels(); sub = g.get_sublabels()
x = case["x"].astype(np.float64)
if x.shape[0] == 32 and x.shape[1] != 32: x = x.T
# first 512 points of key 0 in perm order unknown; print the whole-key S for scale
m = (lab == 1) & (sub == 1)
print("count key0", m.sum()); np.set_printoptions(precision=5, linewidth=200)
print((x[m].T @ x[m])[:3, :6])
