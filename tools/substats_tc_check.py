"""Development aid: fused tensor-core sub-label + statistics kernel vs the separate FFMA sub-label kernel
and the separate statistics kernel on the same state and injected uniforms."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dpmm_pkg
from tests.util import make_niw_case, set_params
pkg = dpmm_pkg.load()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
spread = float(sys.argv[3]) if len(sys.argv) > 3 else 2.5
case = make_niw_case(32, K, n, 1, spread=spread)
rng = np.random.default_rng(5)
u_label, u_sub = rng.random(n), rng.random(n)
bits = rng.integers(0, 2, n).astype(np.uint8)
res = {}
for mode in ("0", "1"):
    os.environ["DPMM_SUBSTATS_TC"] = mode
    g = pkg.GpuSweep(case["x"], case["kind"], seed=1)
    g.set_uniforms(u_label, u_sub, bits)
    set_params(g, case)
    g.sample_labels()
    ll = g.debug_loglik(1)
    g.sample_sublabels()
    sub = g.get_sublabels()
    st = g.suff_stats()
    st_r = g.suff_stats([K, 1] if K > 1 else [1])
    res[mode] = (g.get_labels(), ll, sub, st, st_r)
    if mode == "1":
        g.timing_enable(True)
        for _ in range(20):
            g.sample_labels(); g.sample_sublabels(); g.suff_stats(fetch=False)
        g.sync()
        t = g.timing_read()
        print("fused timing", {k: (round(v[0] / max(v[1], 1) * 1e3, 1), v[1]) for k, v in t.items() if v[1]})
    g.close()
l0, ll0, s0, st0, sr0 = res["0"]
l1, ll1, s1, st1, sr1 = res["1"]
print("labels equal", np.array_equal(l0, l1))
fin = np.isfinite(ll0)
print("sub loglik max abs diff", np.abs(ll1[fin] - ll0[fin]).max(), "max |ll|", np.abs(ll0[fin]).max())
bad = np.nonzero(s0 != s1)[0]
print("sub-label mismatches", bad.size, "of", n)
if bad.size:
    d = ll0[bad]
    print("  |rl - rr| at mismatches (first 10)", np.abs(d[:10, 0] - d[:10, 1]), "ll diff there", np.abs(ll1[bad[:10]] - ll0[bad[:10]]).max(axis=1))
if bad.size == 0:
    c0, sx0, sxx0 = st0; c1, sx1, sxx1 = st1
    print("counts equal", np.array_equal(c0, c1))
    print("sum_x max scaled diff", (np.abs(sx1 - sx0) / (np.sqrt(np.maximum(c0, 1))[..., None] * np.sqrt(np.einsum("msii->msi", sxx0)) + 1e-30)).max())
    dd = np.sqrt(np.einsum("msii->msi", sxx0))
    print("sum_xx max diff / sqrt(SiiSjj)", (np.abs(sxx1 - sxx0) / (dd[..., :, None] * dd[..., None, :] + 1e-30)).max())
    print("symmetric", np.array_equal(sxx1, np.swapaxes(sxx1, -1, -2)))
    print("restricted equal to rows of all:", np.array_equal(sr1[2][0], sxx1[K - 1]), np.array_equal(sr1[0][1], c1[0]))
else:
    # statistics under different sub-labels differ by the mismatching points; compare against Float64 sums of the fused run's own sub-labels
    x = case["x"].astype(np.float64)
    c1, sx1, sxx1 = st1
    worst = 0.0
    for k in range(K):
        for s in (1, 2):
            m = (l1 == k + 1) & (s1 == s)
            S = x[:, m] @ x[:, m].T
            dd = np.sqrt(np.maximum(np.diag(S), 1e-300))
            worst = max(worst, (np.abs(sxx1[k, s] - S) / (dd[:, None] * dd[None, :])).max())
            assert c1[k, s] == m.sum(), (k, s, c1[k, s], m.sum())
    print("fused stats vs Float64 (own sub-labels): worst scaled err", worst)
# worst (cluster, side) of the fused run against Float64 sums of its own sub-labels
x = case["x"].astype(np.float64)
c1, sx1, sxx1 = st1
rows = []
for k in range(K):
    for s in (1, 2):
        m = (l1 == k + 1) & (s1 == s)
        if not m.any():
            continue
        S = x[:, m] @ x[:, m].T
        dd = np.sqrt(np.maximum(np.diag(S), 1e-300))
        E = np.abs(sxx1[k, s] - S) / (dd[:, None] * dd[None, :])
        ex = np.abs(sx1[k, s] - x[:, m].sum(1)) / (np.sqrt(m.sum()) * dd)
        rows.append((E.max(), ex.max(), k, s, int(m.sum()), np.unravel_index(E.argmax(), E.shape)))
rows.sort(reverse=True)
print("worst (err_S, err_x, k, side, N, ij):", rows[:5])
