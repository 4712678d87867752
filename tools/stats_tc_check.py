"""Development aid: tensor-core sufficient statistics vs the FFMA kernel and the Float64 sums."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dpmm_pkg
from tests.util import make_niw_case, set_params
pkg = dpmm_pkg.load()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
case = make_niw_case(32, K, n, 1)
g = pkg.GpuSweep(case["x"], case["kind"], seed=1)
set_params(g, case)
g.sample_labels(); g.sample_sublabels()
os.environ["DPMM_STATS_TC"] = "1"
c1, sx1, sxx1 = g.suff_stats()
os.environ["DPMM_STATS_TC"] = "0"
c0, sx0, sxx0 = g.suff_stats()
print("counts equal", np.array_equal(c0, c1))
print("sum_x max rel diff", np.abs(sx1 - sx0).max() / np.abs(sx0).max())
d = np.sqrt(np.einsum("msii->msi", sxx0))
scale = d[..., :, None] * d[..., None, :] + 1e-30
print("sum_xx max diff / sqrt(SiiSjj)", (np.abs(sxx1 - sxx0) / scale).max())
print("symmetric", np.array_equal(sxx1, np.swapaxes(sxx1, -1, -2)))
# Float64 truth for one key
lab = g.get_labels(); sub = g.get_sublabels()
x = case["x"].astype(np.float64)
if x.shape[0] == 32 and x.shape[1] != 32: x = x.T
np.set_printoptions(precision=4, linewidth=200)
print("tc\n", sxx1[0, 1][:6, :6]); print("ffma\n", sxx0[0, 1][:6, :6])
k = 0
m = (lab == 1) & (sub == 1)
S = x[m].T @ x[m]
dd = np.sqrt(np.diag(S)); sc = dd[:, None] * dd[None, :]
print("vs f64 (key 0 left): tc", (np.abs(sxx1[0, 1] - S) / sc).max(), "ffma", (np.abs(sxx0[0, 1] - S) / sc).max())
os.environ["DPMM_STATS_TC"] = "1"
g.timing_enable(True)
for _ in range(20): g.suff_stats(fetch=False)
t = g.timing_read()
print({k: (round(v[0] / max(v[1], 1) * 1e3, 1), v[1]) for k, v in t.items() if v[1]})
os.environ["DPMM_STATS_TC"] = "0"
for _ in range(20): g.suff_stats(fetch=False)
t = g.timing_read()
print("ffma", {k: (round(v[0] / max(v[1], 1) * 1e3, 1), v[1]) for k, v in t.items() if v[1]})
