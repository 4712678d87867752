"""CPU baseline harness -- TEST/BENCH INFRASTRUCTURE (see oracle/dpmm_oracle.py header).

Runs the restated reference sweep (oracle.OracleSweep: per-cluster passes, materialised n x K
Float32 matrix, Float64 statistics) the way the reference deploys it: W worker PROCESSES over
contiguous column shards (DistributedArrays.distribute layout, dp-parallel-sampling.jl:42-50), one
BLAS thread each (the reference's BLAS.set_num_threads(1) advice, README.md:43, docs/src/perf.md:6),
parameters broadcast to every worker, statistics summed on the master
(local_clusters_actions.jl:171-254).  This is "restated reference (NumPy/OpenBLAS), not Julia".
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

from . import dpmm_oracle as O

_G = {}


def _worker_init():
    try:
        from threadpoolctl import threadpool_limits
        _G["_limit"] = threadpool_limits(1)
    except Exception:
        pass
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    _G["sweeps"] = {}


def _set_params(sw, case):
    if case["kind"] == O.NIW:
        sw.set_params_niw(case["mu"], case["inv_sigma"], case["logdet"], case["weights"], case["lr_weights"])
    else:
        sw.set_params_multinomial(case["log_p"], case["weights"], case["lr_weights"])


def _worker_step(shard):
    lo, hi, seed = shard
    case = _G["case"]
    sw = _G["sweeps"].get((lo, hi))
    if sw is None:
        sw = O.OracleSweep(_G["x"][:, lo:hi], case["kind"], seed=seed, global_offset=lo)
        _G["sweeps"][(lo, hi)] = sw
    _set_params(sw, case)              # broadcast_cluster_params
    sw.sample_labels(False)            # sample_labels_worker!
    sw.sample_sublabels()              # sample_sub_clusters_worker!
    return sw.suff_stats()             # create_suff_stats_dict_worker


class CpuSweepPool:
    """W persistent worker processes (fork), each owning one contiguous shard of the points."""

    def __init__(self, x, case, workers, seed=0):
        self.n = x.shape[1]
        # a worker process is only worth its dispatch cost with a few thousand points to chew on
        self.workers = max(1, min(int(workers), self.n // 5000 if self.n >= 5000 else 1))
        _G["x"] = x
        _G["case"] = case
        bounds = np.linspace(0, self.n, self.workers + 1).astype(np.int64)
        self.shards = [(int(bounds[i]), int(bounds[i + 1]), seed) for i in range(self.workers) if bounds[i + 1] > bounds[i]]
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.workers, initializer=_worker_init)

    def step(self):
        """One sweep: returns the aggregated (counts, sum_x, sum_xx)."""
        parts = self.pool.map(_worker_step, self.shards, chunksize=1)
        counts = sum(p[0] for p in parts)
        sum_x = sum(p[1] for p in parts)
        sum_xx = None if parts[0][2] is None else sum(p[2] for p in parts)   # aggregate_suff_stats
        return counts, sum_x, sum_xx

    def close(self):
        self.pool.close()
        self.pool.join()


def subsample_columns(x, n_sample):
    """Strided sub-sample (generate_gaussian_data lays the clusters out as contiguous blocks, so a
    prefix would not contain the whole mixture)."""
    n = x.shape[1]
    if n_sample >= n:
        return x
    idx = np.linspace(0, n - 1, n_sample).astype(np.int64)
    return np.ascontiguousarray(x[:, idx])


def time_cpu_sweep(x, case, workers, steps, warmup, target_step_s=1.5, min_sample=20000):
    """Times `steps` sweeps on a bounded strided sample sized so that one step takes about
    `target_step_s`.  Returns dict(step_s=[...], n_sample=..., workers=...)."""
    n = x.shape[1]
    probe_n = min(n, max(min_sample, 2000 * workers))
    xs = subsample_columns(x, probe_n)
    pool = CpuSweepPool(xs, case, workers)
    pool.step()
    t0 = time.perf_counter()
    pool.step()
    t_probe = time.perf_counter() - t0
    pool.close()
    n_sample = int(min(n, max(probe_n, probe_n * target_step_s / max(t_probe, 1e-6))))
    xs = subsample_columns(x, n_sample)
    pool = CpuSweepPool(xs, case, workers)
    for _ in range(max(warmup, 1)):
        pool.step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        pool.step()
        ts.append(time.perf_counter() - t0)
    pool.close()
    return dict(step_s=ts, n_sample=n_sample, workers=pool.workers)
