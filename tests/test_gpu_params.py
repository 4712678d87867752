"""Device-side parameter step on the GPU, through the C ABI:
deterministic parts (posterior, log marginal likelihoods, merge table) against the NumPy prior plugin
(dpmmsubclusters.jl_b200/priors.py, which tests/test_host.py pins to the reference's checkpoint bytes);
random parts (Bartlett InverseWishart / mean / Dirichlet draws) against their known moments; and the complete
fit() against the host-side parameter path over seeds."""
import numpy as np
import pytest

import dpmm_pkg
from oracle import dpmm_oracle as O
from tests.util import check_loglik, make_niw_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g
    g.build()
    return dpmm_pkg.load()


def _setup(pkg, D, K, n, seed, alpha=10.0):
    from dpmmsubclusters_jl_b200 import priors as P
    case = make_niw_case(D, K, n, seed=seed, spread=4.0)
    rng = np.random.default_rng(seed)
    lab = rng.integers(1, K + 1, n)
    sub = rng.integers(1, 3, n)
    hyper = P.niw_hyperparams(1.5, rng.standard_normal(D) * 0.3, D + 3.5, np.eye(D) * 1.3 + 0.1)
    g = pkg.GpuSweep(case["x"], pkg.NIW, seed=5)
    g.set_labels(lab); g.set_sublabels(sub)
    g.set_hyper_niw(hyper.κ, hyper.m, hyper.ν, hyper.ψ, alpha)
    return case, g, hyper, lab, sub, P


@pytest.mark.parametrize("D,K,n", [(2, 5, 3000), (32, 6, 6000), (64, 3, 5000), (5, 1, 500)])
def test_posterior_logml_and_merge_table_match_the_prior_plugin(pkg, D, K, n):
    case, g, hyper, lab, sub, P = _setup(pkg, D, K, n, seed=D + K)
    counts, logml, merge = g.posterior_step(None, splittable=np.ones(K, bool))
    sc, sx, sxx = g.suff_stats()
    np.testing.assert_array_equal(counts, sc)
    for k in range(K):
        for s in range(3):
            ss = P.make_suff_stats(hyper, sc[k, s], sx[k, s], sxx[k, s])
            want = P.log_marginal_likelihood(hyper, P.calc_posterior(hyper, ss), ss) if sc[k, s] > 0 else 0.0
            assert abs(logml[k, s] - want) <= 1e-9 * max(1.0, abs(want)) + 2e-3, (k, s, logml[k, s], want)
    for i in range(K):
        for j in range(K):
            if i < j and sc[i, 0] > 0 and sc[j, 0] > 0:
                a = P.make_suff_stats(hyper, sc[i, 0], sx[i, 0], sxx[i, 0])
                b = P.make_suff_stats(hyper, sc[j, 0], sx[j, 0], sxx[j, 0])
                ab = P.aggregate_suff_stats(a, b)
                want = P.log_marginal_likelihood(hyper, P.calc_posterior(hyper, ab), ab)
                assert abs(merge[i, j] - want) <= 1e-9 * abs(want) + 2e-3
            elif merge is not None:
                assert np.isnan(merge[i, j])
    # restricted call + table merge: cluster 1 <- {1 + 2, 1, 2}
    if K >= 2:
        c2, l2, _ = g.posterior_step([2, 1])
        np.testing.assert_array_equal(c2, counts[[1, 0]])
        # (D = 32: the full call is served by the fused tensor-core statistics, the restricted one by the FP32/FP64
        #  kernel; they agree to ~1e-6 relative, and so do the log marginal likelihoods)
        np.testing.assert_allclose(l2, logml[[1, 0]], rtol=5e-6, atol=1e-6)
        g.params_merge(1, 2)
        c3, l3, _ = g.posterior_step([1], from_table=True)
        np.testing.assert_array_equal(c3[0], [counts[0, 0] + counts[1, 0], counts[0, 0], counts[1, 0]])
        np.testing.assert_allclose(l3[0], [merge[0, 1], logml[0, 0], logml[1, 0]], rtol=5e-6, atol=1e-6)
    g.close()


def test_device_posterior_on_the_reference_checkpoint(pkg, golden_dir):
    """The device posterior step on the labels / sub-labels of the reference's checkpoint__50.jld2: the log marginal
    likelihoods must equal those of the posteriors the reference itself stored in that file (post_m / post_psi,
    tests/golden/make_golden.py), evaluated with niw.jl:53-62."""
    import os
    from dpmmsubclusters_jl_b200 import priors as P
    gd = np.load(os.path.join(golden_dir, "niw_2d1k_checkpoint50.npz"))
    hyper = P.niw_hyperparams(gd["prior"][0], np.zeros(2), gd["prior"][1], np.eye(2))
    g = pkg.GpuSweep(gd["x"].astype(np.float32), pkg.NIW)
    g.set_labels(gd["labels"]); g.set_sublabels(gd["sublabels"])
    g.set_hyper_niw(hyper.κ, hyper.m, hyper.ν, hyper.ψ, 100000.0)
    counts, logml, _ = g.posterior_step(None)
    np.testing.assert_array_equal(counts, gd["counts"])
    for k in range(5):
        for s in range(3):
            N = float(gd["counts"][k, s])
            stored = P.niw_hyperparams(1.0 + N, gd["post_m"][k, s], 5.0 + N, gd["post_psi"][k, s])
            ss = P.make_suff_stats(hyper, N, gd["sum_x"][k, s], gd["sum_xx"][k, s])
            want = P.log_marginal_likelihood(hyper, stored, ss)
            # (the device sees the points rounded to Float32, the checkpoint was written from Float64 points: the
            #  statistics differ by ~1e-7 relative, which nu'/2 ~ 100 amplifies in the log determinant term)
            assert abs(logml[k, s] - want) <= 1e-5 * abs(want) + 1e-3, (k, s, logml[k, s], want)
    g.close()


@pytest.mark.parametrize("D", [3, 32])
def test_device_draws_have_the_right_moments_and_feed_the_sweep(pkg, D):
    """Sigma ~ IW(nu', nu' psi')  =>  E[invSigma] = psi'^-1;  mu | Sigma ~ N(m', Sigma / kappa');
    weights ~ Dir(N_1..N_K, alpha);  lr ~ Dir(N_l + a/2, N_r + a/2)."""
    K, n, alpha = 3, 4000, 10.0
    case, g, hyper, lab, sub, P = _setup(pkg, D, K, n, seed=7, alpha=alpha)
    counts, logml, _ = g.posterior_step(None)
    sc, sx, sxx = g.suff_stats()
    R = 300
    inv_acc = np.zeros((K, 3, D, D)); mu_acc = np.zeros((K, 3, D)); w_acc = np.zeros(K); lr_acc = np.zeros((K, 2))
    z_acc = np.zeros((K, 3, D)); z2_acc = np.zeros((K, 3, D))
    posts = [[P.calc_posterior(hyper, P.make_suff_stats(hyper, sc[k, s], sx[k, s], sxx[k, s])) for s in range(3)] for k in range(K)]
    for r in range(R):
        g.sample_params(K)
        mu, lf, ld, w, lr = g.get_params_niw(K)
        L = np.tril(lf)
        inv = L @ np.swapaxes(L, -1, -2)
        inv_acc += inv; mu_acc += mu; w_acc += w; lr_acc += lr
        np.testing.assert_allclose(ld, -np.linalg.slogdet(inv)[1], rtol=2e-5, atol=1e-4)
        for k in range(K):
            for s in range(3):
                # whitened mean residual: sqrt(kappa') L' (mu - m') ~ N(0, I)
                zz = np.sqrt(posts[k][s].κ) * (L[k, s].T @ (mu[k, s] - posts[k][s].m))
                z_acc[k, s] += zz; z2_acc[k, s] += zz * zz
    for k in range(K):
        for s in range(3):
            want = np.linalg.inv(posts[k][s].ψ)
            got = inv_acc[k, s] / R
            scale = np.sqrt(np.outer(np.diag(want), np.diag(want)))
            # relative MC error of a Wishart mean entry ~ sqrt(2 / (nu' R))
            assert np.abs(got - want).max() / scale.max() < 6 * np.sqrt(2.0 / (posts[k][s].ν * R)) + 1e-3
            assert np.abs(z_acc[k, s] / R).max() < 5 / np.sqrt(R)
            assert np.abs(z2_acc[k, s] / R - 1).max() < 6 * np.sqrt(2.0 / R)
    Ntot = sc[:, 0].sum() + alpha
    np.testing.assert_allclose(w_acc / R, sc[:, 0] / Ntot, atol=6 * np.sqrt(0.25 / Ntot / R) + 1e-4)
    want_lr = (sc[:, 1] + alpha / 2) / (sc[:, 1] + sc[:, 2] + alpha)
    np.testing.assert_allclose(lr_acc[:, 0] / R, want_lr, atol=0.01)
    # the packed parameters the sweep reads == the fetched ones: log-likelihood dump vs the oracle on the fetched values
    o = O.OracleSweep(case["x"], O.NIW, seed=5)
    o.set_params_niw(mu, inv.astype(np.float32), ld, w, lr)
    check_loglik(g.debug_loglik(0), o.debug_loglik(0), "device-sampled parameters, label log-likelihood")
    g.close()


def test_fit_device_and_host_parameter_paths_agree_over_seeds(pkg):
    """C1-like data (N=1e4, D=2, K=6) and a 1e5-point D=32 slice of C2: final K and NMI of fit() with the
    parameter step on the device vs on the host (NumPy), 10 seeds each."""
    from dpmmsubclusters_jl_b200 import host as H
    for (N, D, K, iters, burn) in [(10 ** 4, 2, 6, 100, 10), (10 ** 5, 32, 20, 60, 10)]:
        x, labels, _, _ = pkg.generate_gaussian_data(N, D, K, 100.0, np.random.default_rng(5))
        res = {}
        for mode in (True, False):
            nmi, ks = [], []
            for seed in range(10 if D == 2 else 4):
                out = H.fit(x, 10.0, iters=iters, seed=seed, burnout=burn, device_params=mode)
                nmi.append(H.normalized_mutual_info(labels, out[0])); ks.append(len(out[1]))
            res[mode] = (np.array(nmi), np.array(ks))
        print(f"N={N} D={D}: device NMI {np.round(res[True][0], 3)} K {res[True][1]}; host NMI {np.round(res[False][0], 3)} K {res[False][1]}")
        assert abs(res[True][0].mean() - res[False][0].mean()) < 0.05
        assert abs(res[True][1].mean() - res[False][1].mean()) <= 1.5


def test_predict_on_device_matches_the_numpy_posterior_predictive(pkg):
    """predict() runs the Student-t posterior predictive on the GPU (dpmm_predict_niw); labels and
    probabilities against the NumPy formulas of priors.posterior_predictive (niw.jl:68-76) on new points."""
    from dpmmsubclusters_jl_b200 import host as H
    for D, K, N in [(2, 6, 20000), (32, 8, 30000)]:
        x, labels, _, _ = pkg.generate_gaussian_data(N, D, K, 100.0, np.random.default_rng(3))
        out = H.fit(x, 10.0, iters=40, seed=2, burnout=5)
        model = out[-1]
        xnew, _, _, _ = pkg.generate_gaussian_data(5000, D, K, 100.0, np.random.default_rng(4))
        xnew = np.concatenate([xnew, x[:, :3000]], axis=1)
        lg, pg = H.predict(model, xnew, on_device=True)
        lh, ph = H.predict(model, xnew, on_device=False)
        agree = (lg == lh).mean()
        assert agree > 0.999, agree
        np.testing.assert_allclose(pg, ph, atol=2e-3)
        np.testing.assert_allclose(pg.sum(1), 1.0, atol=1e-5)
        # on the training points the posterior-predictive argmax mostly agrees with the sampled labels
        lt, _ = H.predict(model, x, on_device=True)
        print(f"D={D}: device vs NumPy predict agree on {agree:.5f}; predict == fit labels on {(lt == out[0]).mean():.4f}")
        assert (lt == out[0]).mean() > 0.95
