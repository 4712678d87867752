"""Synthetic data of the reference's own generators (the BASELINE configs are defined by them).

  generate_gaussian_data(N, D, K, MixtureVar)   src/data_generators.jl:19-42
  generate_mnmm_data(N, D, K, trials)           src/data_generators.jl:59-72
Same distributions and the same layout (D x N Float32; Gaussian labels are contiguous blocks in
cluster order, :30-39), drawn from a numpy Generator instead of Julia's RNG.
"""
from __future__ import annotations

import numpy as np

from .priors import _inverse_wishart


def generate_gaussian_mixture(D, K, MixtureVar, rng):
    """The mixture part of generate_gaussian_data (data_generators.jl:21, 33-34): weights ~ Dir(1),
    means ~ N(0, MixtureVar I), covariances ~ InverseWishart(D+2, I)."""
    tpi = rng.dirichlet(np.ones(K))
    tmean = np.zeros((D, K), np.float32)
    tcov = np.zeros((D, D, K), np.float32)
    for i in range(K):
        tmean[:, i] = rng.standard_normal(D) * np.sqrt(MixtureVar)
        tcov[:, :, i] = _inverse_wishart(rng, D + 2, np.eye(D))
    return tpi, tmean, tcov


def generate_gaussian_data(N, D, K, MixtureVar, rng=None, shuffle=False, mixture=None):
    """Returns (x [D,N] f32, labels [N] (1-based, Float32 in the reference), means [D,K], covs [D,D,K]).
    `mixture` = (weights, means, covs) reuses a mixture (several shards of one data set)."""
    rng = np.random.default_rng() if rng is None else rng
    tpi, tmean, tcov = generate_gaussian_mixture(D, K, MixtureVar, rng) if mixture is None else mixture
    tzn = rng.multinomial(N, tpi)
    x = np.empty((D, N), np.float32)
    tz = np.zeros(N, np.float32)
    ind = 0
    for i in range(K):
        cnt = int(tzn[i])
        tz[ind:ind + cnt] = i + 1
        if cnt:
            C = tcov[:, :, i].astype(np.float64)
            L = np.linalg.cholesky((C + C.T) / 2)
            x[:, ind:ind + cnt] = (tmean[:, i].astype(np.float64)[:, None]
                                   + L @ rng.standard_normal((D, cnt))).astype(np.float32)
        ind += cnt
    if shuffle:
        p = rng.permutation(N)
        x, tz = np.ascontiguousarray(x[:, p]), tz[p]
    return x, tz, tmean, tcov


def generate_mnmm_data(N, D, K, trials, rng=None):
    """Returns (x [D,N] f32 counts, labels [N] 1-based, clusters [D,K] probability vectors)."""
    rng = np.random.default_rng() if rng is None else rng
    assert K <= D, "the reference indexes alphas[i] for i in 1:K (data_generators.jl:65)"
    clusters = np.zeros((D, K))
    labels = rng.integers(1, K + 1, N)
    for i in range(K):
        alphas = rng.integers(1, 21, D).astype(np.float64)
        alphas[i] = rng.integers(30, 101)
        clusters[:, i] = rng.dirichlet(alphas)
    x = np.empty((D, N), np.float32)
    for i in range(K):
        m = labels == i + 1
        x[:, m] = rng.multinomial(trials, clusters[:, i], size=int(m.sum())).T.astype(np.float32)
    return x, labels, clusters
