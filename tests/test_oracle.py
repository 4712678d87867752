"""CPU tests of the oracle itself: Philox known answers, the reference's golden checkpoints
(stage 3), and hand-checked small cases for stages 1, 2 and 4."""
import os

import numpy as np
import pytest

from oracle import dpmm_oracle as O


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    z = np.uint32(0)
    assert [int(v) for v in O.philox4x32_10(z, z, z, z, 0, 0)] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = np.uint32(0xFFFFFFFF)
    assert [int(v) for v in O.philox4x32_10(f, f, f, f, 0xFFFFFFFF, 0xFFFFFFFF)] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    got = O.philox4x32_10(np.uint32(0x243F6A88), np.uint32(0x85A308D3), np.uint32(0x13198A2E),
                          np.uint32(0x03707344), 0xA4093822, 0x299F31D0)
    assert [int(v) for v in got] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox_uniform_range_and_vectorisation():
    u = O.philox_uniform(1234, O.STREAM_LABEL, 7, np.arange(10000))
    assert u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.02
    u1 = O.philox_uniform(1234, O.STREAM_LABEL, 7, np.array([4321]))
    assert u1[0] == u[4321]
    # 64-bit indices reach the second counter word
    big = O.philox_uniform(1, 0, 1, np.array([2 ** 32 + 5], dtype=np.uint64))
    assert big[0] != O.philox_uniform(1, 0, 1, np.array([5]))[0]


def test_golden_niw_checkpoint(golden_dir):
    g = np.load(os.path.join(golden_dir, "niw_2d1k_checkpoint50.npz"))
    x, labels, sub = g["x"], g["labels"], g["sublabels"]
    for k in range(5):
        m = labels == k + 1
        for s, mask in enumerate([m, m & (sub == 1), m & (sub == 2)]):
            n, sx, S = O.niw_suff_stats(x[:, mask])
            assert n == g["counts"][k, s]
            np.testing.assert_allclose(sx, g["sum_x"][k, s], rtol=1e-10)
            np.testing.assert_allclose(S, g["sum_xx"][k, s], rtol=1e-10)


def test_golden_multinomial_checkpoint_bytes(golden_dir):
    g = np.load(os.path.join(golden_dir, "mnm_1k_checkpoint20.npz"))
    x, labels, sub = g["x"], g["labels"], g["sublabels"]
    d = O.create_suff_stats_dict(x, labels, sub, O.MULTINOMIAL, None, 2)
    for k in range(2):
        for s in range(3):
            n, sx = d[k + 1][s]
            assert n == g["counts"][k, s]
            assert sx.dtype == np.float32
            assert sx.tobytes() == g["sum_x"][k, s].tobytes()      # byte-exact, as stored by the reference


def test_gaussian_loglik_matches_float64_formula_with_d2_constant():
    rng = np.random.default_rng(0)
    D, n = 5, 200
    A = rng.standard_normal((D, D))
    Sigma = A @ A.T + D * np.eye(D)
    mu = rng.standard_normal(D) * 3
    x = (rng.standard_normal((D, n)) * 2 + mu[:, None]).astype(np.float32)
    r = O.gaussian_log_likelihood(x, mu, np.linalg.inv(Sigma), np.linalg.slogdet(Sigma)[1])
    z = x.astype(np.float64) - mu.astype(np.float32).astype(np.float64)[:, None]
    q = np.einsum("ij,ij->j", z, np.linalg.inv(Sigma) @ z)
    ref = -(D * D * np.log(2 * np.pi) + np.linalg.slogdet(Sigma)[1]) / 2 - q / 2   # D^2, not D (G4)
    assert r.dtype == np.float32
    np.testing.assert_allclose(r, ref, rtol=2e-5)


def test_multinomial_loglik():
    rng = np.random.default_rng(1)
    x = rng.integers(0, 9, (7, 30)).astype(np.float32)
    lp = np.log(rng.dirichlet(np.ones(7))).astype(np.float32)
    np.testing.assert_allclose(O.multinomial_log_likelihood(x, lp), lp.astype(np.float64) @ x, rtol=1e-5)


def _scalar_inverse_cdf(p, u):
    """Literal transcription of the StatsBase loop for one row (pure Python, Float32 cw)."""
    s = np.float32(0)
    for v in p:
        s = np.float32(s + v)
    t = float(u) * float(s)
    i, cw = 0, np.float32(p[0])
    while float(cw) < t and i < len(p) - 1:
        i += 1
        cw = np.float32(cw + p[i])
    return i + 1


def test_inverse_cdf_against_scalar_loop():
    rng = np.random.default_rng(2)
    M = (rng.standard_normal((500, 6)) * 4).astype(np.float32)
    M[3, 2] = np.nan
    M[4, :] = -np.inf          # all -Inf row => NaN weights => label 1
    u = rng.random(500)
    u[0], u[1] = 0.0, np.nextafter(1.0, 0.0)
    P = O.softmax_rows_f32(M)
    got = O.sample_log_cat_array(M, u)
    want = np.array([_scalar_inverse_cdf(P[i], u[i]) for i in range(500)])
    np.testing.assert_array_equal(got, want)
    assert got[4] == 1 and got[0] >= 1 and got.max() <= 6
    assert P[3, 2] == 0.0


def test_inverse_cdf_distribution():
    rng = np.random.default_rng(3)
    logp = np.log(np.array([0.1, 0.2, 0.3, 0.4], np.float32))
    M = np.tile(logp, (40000, 1))
    got = O.sample_log_cat_array(M, rng.random(40000))
    freq = np.bincount(got, minlength=5)[1:] / 40000
    np.testing.assert_allclose(freq, [0.1, 0.2, 0.3, 0.4], atol=0.01)


def test_argmax_rows_first_max_and_nan():
    M = np.array([[1, 3, 3, 2], [0, np.nan, 5, np.nan], [-np.inf, -np.inf, -np.inf, -np.inf]], np.float32)
    np.testing.assert_array_equal(O.argmax_rows(M), [2, 2, 1])


def test_relabel_split_merge_remove():
    labels = np.array([1, 1, 2, 2, 2, 3, 3, 1], np.int64)
    sub = np.array([1, 2, 1, 2, 2, 1, 2, 2], np.int64)
    bits = np.array([0, 1, 0, 1, 0, 1, 0, 1], np.int64)
    # split cluster 2 -> new index 4 : sub==2 points move, both halves get fresh bits
    l, s = O.split_cluster_local(labels.copy(), sub.copy(), [2], [4], bits)
    np.testing.assert_array_equal(l, [1, 1, 2, 4, 4, 3, 3, 1])
    np.testing.assert_array_equal(s, [1, 2, 1, 2, 1, 1, 2, 2])
    # merge 1 <- 3
    l2, s2 = O.merge_clusters(l.copy(), s.copy(), [1], [3])
    np.testing.assert_array_equal(l2, [1, 1, 2, 4, 4, 1, 1, 1])
    np.testing.assert_array_equal(s2, [1, 1, 1, 2, 1, 2, 2, 1])
    # cluster 3 is now empty -> compaction
    l3 = O.remove_empty_clusters(l2.copy(), [5, 1, 0, 2])
    np.testing.assert_array_equal(l3, [1, 1, 2, 3, 3, 1, 1, 1])
    # two empties, one of them first
    l4 = O.remove_empty_clusters(np.array([2, 4, 5, 5], np.int64), [0, 1, 0, 1, 2])
    np.testing.assert_array_equal(l4, [1, 2, 3, 3])


def test_reset_bad_clusters_only_touches_listed():
    labels = np.array([1, 2, 2, 3], np.int64)
    sub = np.array([1, 1, 1, 1], np.int64)
    out = O.reset_bad_clusters(labels, sub.copy(), [2], np.array([1, 1, 0, 1]))
    np.testing.assert_array_equal(out, [1, 2, 1, 1])
    out = O.reset_bad_clusters(labels, sub.copy(), None, np.array([1, 1, 0, 1]))
    np.testing.assert_array_equal(out, [2, 2, 1, 2])


def test_sweep_object_end_to_end_small():
    rng = np.random.default_rng(5)
    D, n, K = 3, 400, 3
    mus = rng.standard_normal((K, D)) * 6
    x = np.concatenate([rng.standard_normal((D, n // K + (1 if k == 0 else 0))) + mus[k][:, None] for k in range(K)], axis=1).astype(np.float32)
    n = x.shape[1]
    sw = O.OracleSweep(x, O.NIW, seed=11)
    sw.init_labels(1)
    assert set(np.unique(sw.get_labels())) == {1} and set(np.unique(sw.get_sublabels())) <= {1, 2}
    mu = np.zeros((K, 3, D), np.float32)
    inv = np.tile(np.eye(D, dtype=np.float32), (K, 3, 1, 1))
    for k in range(K):
        mu[k, 0] = mus[k]; mu[k, 1] = mus[k] - 0.5; mu[k, 2] = mus[k] + 0.5
    sw.set_params_niw(mu, inv, np.zeros((K, 3), np.float32), np.full(K, 1 / K, np.float32), np.full((K, 2), 0.5, np.float32))
    sw.sample_labels(final=False)
    sw.sample_sublabels()
    counts, sx, sxx = sw.suff_stats()
    assert counts[:, 0].sum() == n
    np.testing.assert_array_equal(counts[:, 0], counts[:, 1] + counts[:, 2])
    np.testing.assert_allclose(sxx[:, 0], sxx[:, 1] + sxx[:, 2], rtol=1e-9, atol=1e-9)
    # well separated clusters: sampled labels recover the blocks
    lab = sw.get_labels()
    assert (lab[: n // K] == lab[0]).mean() > 0.99
    # restricted statistics only cover the listed clusters
    c2, _, _ = sw.suff_stats([2])
    np.testing.assert_array_equal(c2[0], counts[1])


def test_near_tie_mask():
    logm = np.log(np.array([[0.25, 0.25, 0.5]], np.float64)).astype(np.float32)
    assert O.near_tie_mask(logm, [0.25 + 1e-7], [1], 1e-5)[0]
    assert not O.near_tie_mask(logm, [0.4], [1], 1e-5)[0]
