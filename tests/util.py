"""Shared helpers for the parity tests: synthetic cases and the GPU-vs-oracle comparison with the
near-tie accounting of SURVEY.md 8c."""
import numpy as np

from oracle import dpmm_oracle as O

LOG_TOL = 1e-5          # north_star: labels bit-exact "except documented near-ties within 1e-5 in log space"
LL_RTOL = 1e-4          # north_star: log-likelihoods within 1e-4 relative (fp32)
STATS_RTOL = 1e-4       # north_star: sum x / sum xx' within 1e-4 relative (to sqrt(S_ii S_jj))


def random_spd(rng, D, scale=1.0):
    A = rng.standard_normal((D, D))
    S = A @ A.T / D + 0.5 * np.eye(D)
    return S * scale


def make_niw_case(D, K, n, seed, spread=2.5):
    """Overlapping Gaussian clusters (so that the draws are not all deterministic) with 3K sampled
    distributions in the reference's dtype flow: Float64 draw -> Float32 mu / invSigma / logdet."""
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((K, D)) * spread
    z = rng.integers(0, K, n)
    covs = [random_spd(rng, D) for _ in range(K)]
    x = np.empty((D, n), np.float32)
    for k in range(K):
        m = z == k
        Lc = np.linalg.cholesky(covs[k])
        x[:, m] = (centers[k][:, None] + Lc @ rng.standard_normal((D, int(m.sum())))).astype(np.float32)
    mu = np.zeros((K, 3, D), np.float32)
    inv = np.zeros((K, 3, D, D), np.float32)
    logdet = np.zeros((K, 3), np.float32)
    for k in range(K):
        for s in range(3):
            Sig = covs[k] * (1.0 if s == 0 else 0.8) + 0.05 * random_spd(rng, D)
            m = centers[k] + (0 if s == 0 else (0.7 if s == 1 else -0.7)) * rng.standard_normal(D)
            mu[k, s] = m
            inv[k, s] = np.linalg.inv(Sig)          # Float64 inverse rounded to Float32 (niw.jl:37,39)
            logdet[k, s] = np.linalg.slogdet(Sig)[1]
    w = rng.dirichlet(np.ones(K) * 5).astype(np.float32)
    lr = rng.dirichlet(np.ones(2) * 5, size=K).astype(np.float32)
    return dict(kind=O.NIW, x=x, mu=mu, inv_sigma=inv, logdet=logdet, weights=w, lr_weights=lr, K=K, D=D, n=n)


def make_mnm_case(D, K, n, seed, trials=50):
    rng = np.random.default_rng(seed)
    probs = rng.dirichlet(np.ones(D) * 0.7, size=K)
    z = rng.integers(0, K, n)
    x = np.empty((D, n), np.float32)
    for k in range(K):
        m = z == k
        x[:, m] = rng.multinomial(trials, probs[k], size=int(m.sum())).T.astype(np.float32)
    log_p = np.zeros((K, 3, D), np.float32)
    for k in range(K):
        for s in range(3):
            p = 0.85 * probs[k] + 0.15 * rng.dirichlet(np.ones(D))
            log_p[k, s] = np.log(p)
    w = rng.dirichlet(np.ones(K) * 5).astype(np.float32)
    lr = rng.dirichlet(np.ones(2) * 5, size=K).astype(np.float32)
    return dict(kind=O.MULTINOMIAL, x=x, log_p=log_p, weights=w, lr_weights=lr, K=K, D=D, n=n)


def set_params(sw, case):
    if case["kind"] == O.NIW:
        sw.set_params_niw(case["mu"], case["inv_sigma"], case["logdet"], case["weights"], case["lr_weights"])
    else:
        sw.set_params_multinomial(case["log_p"], case["weights"], case["lr_weights"])


def tie_tolerance(logmat):
    """Log-space band for the near-tie test: the north_star's 1e-5 plus two Float32 ulps of the
    largest log-likelihood in the row (the reference's own values are quantised to that ulp: with
    the D^2 constant |r| ~ 1e3 at D=32, i.e. ulp = 6e-5 > 1e-5)."""
    a = np.abs(np.where(np.isfinite(logmat), logmat, 0)).max(axis=1).astype(np.float32)
    return LOG_TOL + 2.0 * np.spacing(a).astype(np.float64)


def check_draws(logmat, u, got, want, what):
    """`got` (GPU) must equal `want` (oracle) except where the uniform is a documented near-tie."""
    got = np.asarray(got); want = np.asarray(want)
    bad = np.nonzero(got != want)[0]
    if bad.size:
        ok = O.near_tie_mask(logmat[bad], np.asarray(u)[bad], got[bad], tie_tolerance(logmat[bad]))
        assert ok.all(), (f"{what}: {int((~ok).sum())} of {bad.size} mismatches are NOT near-ties; first: "
                          f"i={bad[~ok][0]} got={got[bad[~ok][0]]} want={want[bad[~ok][0]]} "
                          f"row={logmat[bad[~ok][0]]} u={np.asarray(u)[bad[~ok][0]]}")
    assert bad.size <= max(3, 2e-3 * got.size), f"{what}: too many near-tie mismatches ({bad.size}/{got.size})"
    return int(bad.size)


def check_loglik(got, want, what):
    got = np.asarray(got, np.float64); want = np.asarray(want, np.float64)
    fin = np.isfinite(want)
    assert (np.isfinite(got) == fin).all(), f"{what}: finiteness pattern differs"
    err = np.abs(got[fin] - want[fin]) / np.maximum(np.abs(want[fin]), 1.0)
    assert err.max(initial=0.0) <= LL_RTOL, f"{what}: max rel err {err.max():.3e}"
    return float(err.max(initial=0.0))


def check_stats(gpu, ora, prior_kind, what):
    gc, gsx, gsxx = gpu
    oc, osx, osxx = ora
    np.testing.assert_array_equal(gc, oc, err_msg=f"{what}: counts")
    if prior_kind == O.MULTINOMIAL:
        np.testing.assert_array_equal(gsx, osx.astype(np.float64), err_msg=f"{what}: count vectors must be exact")
        return 0.0
    # scale: sqrt(S_ii S_jj) for S, sqrt(N S_ii) for sum x (Cauchy-Schwarz bounds of the entries)
    diag = np.sqrt(np.maximum(np.einsum("msii->msi", osxx), 1e-300))
    e_xx = np.abs(gsxx - osxx) / np.maximum(diag[..., :, None] * diag[..., None, :], 1e-30)
    e_x = np.abs(gsx - osx) / np.maximum(np.sqrt(np.maximum(oc, 1))[..., None] * diag, 1e-30)
    nz = oc > 0
    worst = max(e_xx[nz].max(initial=0.0), e_x[nz].max(initial=0.0))
    assert worst <= STATS_RTOL, f"{what}: max scaled err {worst:.3e}"
    assert np.abs(gsxx[~nz]).max(initial=0.0) == 0 and np.abs(gsx[~nz]).max(initial=0.0) == 0, f"{what}: empty sets must give zeros"
    np.testing.assert_array_equal(gsxx, np.swapaxes(gsxx, -1, -2), err_msg=f"{what}: S must be symmetric")
    return float(worst)


def compare_sweeps(g, o, case, rng, final=False, warm=False):
    """Drive the GPU sweep `g` and the oracle `o` through one full iteration on the same injected
    randomness and compare every stage.  After each stage the GPU state is copied into the oracle so
    that a near-tie at one stage cannot cascade into the next comparison.
    warm: first give both sides the labels of an argmax pass, as an iteration deep inside a run would find
    them (the D = 32 / 64 tensor-core label kernel walks the points in the order of the current labels and
    uses each tile's old cluster as its pivot; cold = every point starts in cluster 1)."""
    n, K = case["n"], case["K"]
    u_label, u_sub = rng.random(n), rng.random(n)
    bits = rng.integers(0, 2, n).astype(np.uint8)
    for s in (g, o):
        s.set_uniforms(u_label, u_sub, bits)
        set_params(s, case)
    if warm:
        g.sample_labels(True); o.sample_labels(True)
        o.set_labels(g.get_labels())
    rep = {}
    # stage 1: log-likelihood matrices (labels phase)
    LLo = o.debug_loglik(0)
    rep["ll_err"] = check_loglik(g.debug_loglik(0), LLo, "label log-likelihood")
    # stage 2a: labels
    g.sample_labels(final); o.sample_labels(final)
    gl, ol = g.get_labels(), o.get_labels()
    if final:
        bad = np.nonzero(gl != ol)[0]
        if bad.size:   # (two different labels exist, so K >= 2)
            srt = np.sort(LLo[bad], axis=1)
            assert ((srt[:, -1] - srt[:, -2]) <= tie_tolerance(LLo[bad])).all(), "argmax mismatch that is not a tie"
        rep["label_ties"] = int(bad.size)
    else:
        rep["label_ties"] = check_draws(LLo, u_label, gl, ol, "labels")
    o.set_labels(gl)
    # stage 2b: sub-labels under the (GPU's) fresh labels
    SLo = o.debug_loglik(1)
    rep["sub_ll_err"] = check_loglik(g.debug_loglik(1), SLo, "sub-label log-likelihood")
    g.sample_sublabels(); o.sample_sublabels()
    gs, os_ = g.get_sublabels(), o.get_sublabels()
    rep["sub_ties"] = check_draws(SLo, u_sub, gs, os_, "sub-labels")
    o.set_sublabels(gs)
    np.testing.assert_array_equal(g.get_labels(), gl, err_msg="sub-label sampling must not touch labels")
    # stage 3: statistics, all clusters and a restricted subset
    rep["stats_err"] = check_stats(g.suff_stats(), o.suff_stats(), case["kind"], "suff stats (all)")
    sub_idx = [K, 1] if K > 1 else [1]
    check_stats(g.suff_stats(sub_idx), o.suff_stats(sub_idx), case["kind"], "suff stats (restricted)")
    # stage 4: split the biggest cluster into a new index, merge two, compact
    counts = np.bincount(gl, minlength=K + 1)[1:]
    big = int(np.argmax(counts)) + 1
    for s in (g, o):
        s.apply_split([big], [K + 1])
    np.testing.assert_array_equal(g.get_labels(), o.get_labels(), err_msg="split: labels")
    np.testing.assert_array_equal(g.get_sublabels(), o.get_sublabels(), err_msg="split: sub-labels")
    idx = [big, K + 1]
    for s in (g, o):
        s.K = K + 1  # both mirrors now hold K+1 clusters (host appended one)
    check_stats(g.suff_stats(idx), o.suff_stats(idx), case["kind"], "suff stats after split")
    if K >= 2:
        i, j = (1 if big != 1 else 2), K + 1
        for s in (g, o):
            s.apply_merge([i], [j])
        np.testing.assert_array_equal(g.get_labels(), o.get_labels(), err_msg="merge: labels")
        np.testing.assert_array_equal(g.get_sublabels(), o.get_sublabels(), err_msg="merge: sub-labels")
    cnt = np.bincount(g.get_labels(), minlength=K + 2)[1:K + 2]
    for s in (g, o):
        s.remove_empty(cnt)
    np.testing.assert_array_equal(g.get_labels(), o.get_labels(), err_msg="remove_empty: labels")
    for s in (g, o):
        s.randomize_sublabels([1])
    np.testing.assert_array_equal(g.get_sublabels(), o.get_sublabels(), err_msg="randomize_sublabels(indices)")
    for s in (g, o):
        s.randomize_sublabels(None)
    np.testing.assert_array_equal(g.get_sublabels(), o.get_sublabels(), err_msg="randomize_sublabels(all)")
    for s in (g, o):
        s.set_uniforms(None, None, None)
    return rep
