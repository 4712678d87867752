"""End-to-end fit() on the GPU (C ABI underneath): the reference's own test/module_tests.jl testsets,
the C1 example of docs/src/getting_started.md, and statistical equivalence with the same host logic
running on the oracle (NMI / final K over seeds)."""
import numpy as np
import pytest

import dpmm_pkg
from oracle import dpmm_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g
    g.build()
    return dpmm_pkg.load()


def oracle_factory(x, kind, seed, goff):
    return O.OracleSweep(x, kind, seed=seed, global_offset=goff)


def test_module_deterministic_four_point_masses(pkg):
    """test/module_tests.jl:10-32."""
    x = np.zeros((2, 1000), np.float32)
    x[:, 0:250] = [[-1], [-1]]; x[:, 250:500] = [[-1], [1]]; x[:, 500:750] = [[1], [-1]]; x[:, 750:1000] = [[1], [1]]
    labels, clusters, weights, *_r, dp_model = pkg.fit(x, 100.0, iters=200, seed=123456789, burnout=15)
    assert len(clusters) == 4
    assert all(w >= 0.15 for w in weights)
    lbls, _ = pkg.predict(dp_model, x)
    np.testing.assert_array_equal(lbls, labels)
    assert [c for _, c in pkg.get_labels_histogram(labels)] == [250, 250, 250, 250]


def test_module_random_mess(pkg):
    """test/module_tests.jl:36-47, full size."""
    x, labels, _, _ = pkg.generate_gaussian_data(10 ** 5, 3, 10, 100.0, np.random.default_rng(0))
    hyper = pkg.niw_hyperparams(1.0, np.zeros(3), 5, np.eye(3))
    out = pkg.fit(x, hyper, 1e21, iters=100, seed=12345, gt=labels)
    assert len(out[1]) > 1
    assert out[4][-1] > 0.6


def test_module_multinomial(pkg):
    """test/module_tests.jl:49-60 (without save/load)."""
    x, labels, _ = pkg.generate_mnmm_data(10 ** 3, 100, 20, 50, np.random.default_rng(0))
    hyper = pkg.multinomial_hyper(np.ones(100, np.float32))
    out = pkg.fit(x, hyper, 1e5, iters=39, seed=3, gt=labels)
    assert len(out[1]) > 1


def test_c1_getting_started_example(pkg):
    """BASELINE config C1 / docs/src/getting_started.md:27-37: N=1e4, D=2, K=6, alpha=10, 100 iterations; the
    documented run ends at K=6 with NMI 1.0.  Five sampler seeds on one data set, both parameter paths: the
    documented end state must be the typical one."""
    from dpmmsubclusters_jl_b200 import host as H
    x, labels, _, _ = pkg.generate_gaussian_data(10 ** 4, 2, 6, 100.0, np.random.default_rng(5))
    k_true = len(np.unique(labels))
    for device_params in (True, False):
        ks, nmis = [], []
        for seed in range(5):
            out = H.fit(x, 10.0, iters=100, seed=seed, gt=labels, burnout=10, device_params=device_params)
            ks.append(len(out[1])); nmis.append(out[4][-1])
        print(f"C1 device_params={device_params}: final K {ks} (true {k_true}), NMI {np.round(nmis, 4)}")
        assert int(np.median(ks)) == k_true and max(abs(k - k_true) for k in ks) <= 1
        assert np.mean(nmis) > 0.98


def test_gpu_and_oracle_hosts_statistically_indistinguishable(pkg):
    """Same host logic, same seeds, workers = GPU vs oracle: NMI against the ground truth and the
    final K agree across 10 seeds (they share the Philox streams, so most runs coincide exactly)."""
    from dpmmsubclusters_jl_b200 import host as H
    x, labels, _, _ = pkg.generate_gaussian_data(3000, 2, 4, 100.0, np.random.default_rng(11))
    nmi_g, nmi_o, k_g, k_o, same = [], [], [], [], 0
    for seed in range(10):
        g = H.fit(x, 10.0, iters=40, seed=seed, gt=labels, burnout=5, device_params=False)
        o = H.fit(x, 10.0, iters=40, seed=seed, gt=labels, burnout=5, sweep_factory=oracle_factory)
        nmi_g.append(g[4][-1]); nmi_o.append(o[4][-1]); k_g.append(len(g[1])); k_o.append(len(o[1]))
        same += int(np.array_equal(g[0], o[0]))
    print("NMI gpu", np.round(nmi_g, 3), "oracle", np.round(nmi_o, 3), "K gpu", k_g, "oracle", k_o, "identical runs", same)
    assert abs(np.mean(nmi_g) - np.mean(nmi_o)) < 0.05
    assert abs(np.mean(k_g) - np.mean(k_o)) <= 1.0
    assert same >= 5


@pytest.mark.timeout(300, method="thread")
def test_full_size_c2_fit_from_one_cluster(pkg):
    """fit() on the full C2 shape (N = 1e6, D = 32, K_true = 20) from K = 1, twice: every shape of the fused sub-label +
    statistics kernel's tile sequence on the way (one huge cluster, freshly split clusters, tiny and empty ones), with
    its three epilogue groups free to drift apart -- the run has to finish (no barrier phase may be skipped) and land
    where the host-parameter path lands."""
    x, z, _, _ = pkg.generate_gaussian_data(1_000_000, 32, 20, 100.0, np.random.default_rng(0))
    for seed in (1, 2):
        out = pkg.fit(x, 10.0, iters=100, seed=seed, burnout=20)
        from dpmmsubclusters_jl_b200.host import normalized_mutual_info
        nmi = normalized_mutual_info(z, out[0])
        assert 10 <= len(out[1]) <= 24 and nmi > 0.9, (len(out[1]), nmi)


@pytest.mark.timeout(300, method="thread")
def test_d64_fit_from_one_cluster(pkg):
    """fit() at D = 64 from K = 1: the tcgen05 label, sub-label and statistics kernels of the C5 shape through every tile
    sequence a run produces (one cluster, fresh splits, tiny and empty clusters), on both parameter paths."""
    x, z, _, _ = pkg.generate_gaussian_data(300_000, 64, 10, 100.0, np.random.default_rng(4))
    from dpmmsubclusters_jl_b200.host import normalized_mutual_info
    for device_params in (True, False):
        out = pkg.fit(x, 10.0, iters=60 if device_params else 40, seed=3, burnout=10, device_params=device_params)
        nmi = normalized_mutual_info(z, out[0])
        assert 5 <= len(out[1]) <= 14 and nmi > 0.85, (device_params, len(out[1]), nmi)
