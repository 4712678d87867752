"""Host logic (the Python mirror of the reference's master functions) on CPU, with the oracle standing
in for the workers.  The three testsets are the reference's own test/module_tests.jl, ported."""
import numpy as np
import pytest

import dpmm_pkg
from oracle import dpmm_oracle as O

pkg = dpmm_pkg.load()
from dpmmsubclusters_jl_b200 import host as H  # noqa: E402
from dpmmsubclusters_jl_b200 import priors as P  # noqa: E402


def oracle_factory(x, kind, seed, goff):
    return O.OracleSweep(x, kind, seed=seed, global_offset=goff)


def four_point_masses():
    """test/module_tests.jl:1-8: 1000 points = 4 point masses (+-1, +-1) x 250."""
    x = np.zeros((2, 1000), np.float32)
    x[:, 0:250] = [[-1], [-1]]
    x[:, 250:500] = [[-1], [1]]
    x[:, 500:750] = [[1], [-1]]
    x[:, 750:1000] = [[1], [1]]
    return x


def test_niw_posterior_and_marginal_closed_forms():
    rng = np.random.default_rng(0)
    D, n = 3, 50
    pts = rng.standard_normal((D, n)) + 2
    prior = P.niw_hyperparams(1.0, np.zeros(D), D + 3, np.eye(D))
    ss = P.make_suff_stats(prior, n, pts.sum(1), pts @ pts.T)
    post = P.calc_posterior(prior, ss)
    assert post.κ == 1 + n and post.ν == D + 3 + n
    np.testing.assert_allclose(post.m, pts.sum(1) / (1 + n))
    # psi' nu' = nu psi + sum (x - xbar)(x - xbar)' + kappa n/(kappa+n) xbar xbar'   (m0 = 0)
    xbar = pts.mean(1)
    C = (pts - xbar[:, None]) @ (pts - xbar[:, None]).T
    want = (prior.ν * prior.ψ + C + (n / (1 + n)) * np.outer(xbar, xbar)) / post.ν
    np.testing.assert_allclose(post.ψ, want, rtol=1e-10)
    assert P.calc_posterior(prior, P.empty_suff_stats(prior)) is prior
    # marginal likelihood is additive-consistent: p(A u B) = p(A) p(B | A)
    a, b = pts[:, :20], pts[:, 20:]
    ssa = P.make_suff_stats(prior, 20, a.sum(1), a @ a.T)
    posta = P.calc_posterior(prior, ssa)
    ssb = P.make_suff_stats(prior, n - 20, b.sum(1), b @ b.T)
    lhs = P.log_marginal_likelihood(prior, post, ss)
    rhs = P.log_marginal_likelihood(prior, posta, ssa) + P.log_marginal_likelihood(posta, P.calc_posterior(posta, ssb), ssb)
    assert abs(lhs - rhs) < 1e-3 * max(1, abs(lhs))      # log_multivariate_gamma accumulates in Float32 (utils.jl:66-72)


def test_calc_posterior_matches_the_posteriors_stored_in_the_reference_checkpoints(golden_dir):
    """The reference's own checkpoints hold every cluster's posterior hyper-parameters next to its statistics
    (tests/golden/make_golden.py): m' and psi' of niw.jl:20-31 for checkpoint__50.jld2, alpha' of
    multinomial_prior.jl:16-21 for checkpoint_20.jld2.  calc_posterior must reproduce the stored bytes."""
    import os
    gd = np.load(os.path.join(golden_dir, "niw_2d1k_checkpoint50.npz"))
    hyper = P.niw_hyperparams(gd["prior"][0], np.zeros(2), gd["prior"][1], np.eye(2))
    for k in range(5):
        for s in range(3):
            ss = P.make_suff_stats(hyper, gd["counts"][k, s], gd["sum_x"][k, s], gd["sum_xx"][k, s])
            post = P.calc_posterior(hyper, ss)
            np.testing.assert_allclose(post.m, gd["post_m"][k, s], rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(post.ψ, gd["post_psi"][k, s], rtol=1e-11, atol=1e-13)
            assert post.κ == 1.0 + gd["counts"][k, s] and post.ν == 5.0 + gd["counts"][k, s]
    gm = np.load(os.path.join(golden_dir, "mnm_1k_checkpoint20.npz"))
    mh = P.multinomial_hyper(np.ones(100, np.float32))
    for k in range(2):
        for s in range(3):
            ss = P.make_suff_stats(mh, gm["counts"][k, s], gm["sum_x"][k, s])
            assert P.calc_posterior(mh, ss).α.astype(np.float32).tobytes() == gm["post_alpha"][k, s].tobytes()
    # generate_mnmm_data(N, D, K, trials) invariants of the reference's own data file: every point is 50 draws
    assert (gm["row_sums"] == 50).all() and np.array_equal(gm["x"], np.round(gm["x"])) and gm["x"].min() >= 0


def test_sample_distribution_moments():
    rng = np.random.default_rng(1)
    D = 3
    h = P.niw_hyperparams(50.0, np.array([1.0, -2.0, 0.5]), 200.0, np.diag([1.0, 2.0, 0.5]))
    S = np.mean([P.sample_distribution(h, rng).Σ.astype(np.float64) for _ in range(400)], axis=0)
    # E[Sigma] = nu psi / (nu - D - 1)
    np.testing.assert_allclose(S, h.ν * h.ψ / (h.ν - D - 1), rtol=0.08, atol=0.03)
    d = P.sample_distribution(h, rng)
    np.testing.assert_allclose(d.invΣ.astype(np.float64) @ d.Σ.astype(np.float64), np.eye(D), atol=1e-4)
    assert abs(d.logdetΣ - np.linalg.slogdet(d.Σ.astype(np.float64))[1]) < 1e-4
    m = P.sample_distribution(P.multinomial_hyper(np.ones(5)), rng)
    assert abs(np.exp(m.α.astype(np.float64)).sum() - 1) < 1e-5


def test_module_deterministic_four_point_masses():
    """test/module_tests.jl:10-32 ("Testing Module (Determinstic)")."""
    x = four_point_masses()
    labels, clusters, weights, *_rest, dp_model = H.fit(x, 100.0, iters=200, seed=123456789, burnout=15,
                                                        sweep_factory=oracle_factory)
    assert len(clusters) == 4
    assert all(w >= 0.15 for w in weights)
    lbls, _ = H.predict(dp_model, x)
    np.testing.assert_array_equal(lbls, labels)
    assert [c for _, c in H.get_labels_histogram(labels)] == [250, 250, 250, 250]


def test_module_random_mess():
    """test/module_tests.jl:36-47 (N reduced from 1e5 to 2e4 to keep the CPU suite short; the GPU test runs 1e5)."""
    x, labels, _, _ = pkg.generate_gaussian_data(20000, 3, 10, 100.0, np.random.default_rng(0))
    hyper = P.niw_hyperparams(1.0, np.zeros(3), 5, np.eye(3))
    out = H.fit(x, hyper, 1e21, iters=60, seed=12345, gt=labels, sweep_factory=oracle_factory)
    assert len(out[1]) > 1
    assert out[4][-1] > 0.5            # NMI against the ground truth


def test_module_multinomial():
    """test/module_tests.jl:49-60 without the save/load part (checkpoints are out of scope)."""
    x, labels, _ = pkg.generate_mnmm_data(1000, 100, 20, 50, np.random.default_rng(0))
    hyper = P.multinomial_hyper(np.ones(100, np.float32))
    out = H.fit(x, hyper, 1e5, iters=39, seed=3, gt=labels, sweep_factory=oracle_factory)
    assert len(out[1]) > 1


def test_nmi_matches_definition():
    a = np.array([1, 1, 2, 2, 3, 3])
    assert abs(H.normalized_mutual_info(a, a) - 1) < 1e-12
    assert abs(H.normalized_mutual_info(a, np.array([5, 5, 7, 7, 9, 9])) - 1) < 1e-12
    assert H.normalized_mutual_info(a, np.array([1, 2, 1, 2, 1, 2])) < 1e-9
    try:
        from sklearn.metrics import normalized_mutual_info_score
        b = np.random.default_rng(0).integers(0, 4, 500); c = np.where(np.random.default_rng(1).random(500) < 0.7, b, 3 - b)
        assert abs(H.normalized_mutual_info(b, c) - normalized_mutual_info_score(b, c)) < 1e-9
    except ImportError:
        pass
