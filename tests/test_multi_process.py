"""N>1 host path on CPU: world_size-2 gloo processes, each owning one contiguous shard of the points
(the reference's DistributedArrays layout), the oracle standing in for the per-GPU kernels and a gloo
all-reduce standing in for the NCCL all-reduce of the packed statistics.  Checks what must hold on
the GPU box too: (a) Philox keyed by the GLOBAL point index makes labels independent of the sharding,
(b) all-reduced statistics equal the single-process ones, (c) the SPMD host (same seed on every rank)
takes identical split/merge decisions, so a sharded fit equals the unsharded fit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dpmm_pkg
from oracle import dpmm_oracle as O
from tests.util import make_niw_case, set_params

pkg = dpmm_pkg.load()
from dpmmsubclusters_jl_b200 import host as H  # noqa: E402


class AllReduceSweep(O.OracleSweep):
    """OracleSweep whose suff_stats ends in an all-reduce, like dpmm_suff_stats with a communicator."""

    def suff_stats(self, indices=None):
        c, sx, sxx = super().suff_stats(indices)
        packed = torch.from_numpy(np.concatenate([c.ravel().astype(np.float64), sx.ravel(),
                                                  sxx.ravel() if sxx is not None else np.zeros(0)]))
        dist.all_reduce(packed)
        p = packed.numpy()
        c2 = np.rint(p[:c.size]).astype(np.int64).reshape(c.shape)
        sx2 = p[c.size:c.size + sx.size].reshape(sx.shape)
        sxx2 = p[c.size + sx.size:].reshape(sxx.shape) if sxx is not None else None
        return c2, sx2, sxx2


    # the master's reductions of smart_cluster_init! (local_clusters_actions.jl:578-612), as dpmm_smart_* do them
    def smart_project(self, cluster, v, mu):
        lo, hi, cnt = super().smart_project(cluster, v, mu)
        inf = float("inf")
        t = torch.tensor([-lo if cnt > 1 else -inf, hi if cnt > 1 else -inf], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c = torch.tensor([float(cnt)], dtype=torch.float64)
        dist.all_reduce(c)
        if t[1].item() == -inf:
            return float("nan"), float("nan"), int(c.item())
        return -t[0].item(), t[1].item(), int(c.item())

    def smart_kmeans_iter(self, min_mean, max_mean):
        t = torch.tensor(super().smart_kmeans_iter(min_mean, max_mean), dtype=torch.float64)
        dist.all_reduce(t)
        return tuple(t.tolist())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = make_niw_case(4, 5, 3000, seed=8)
        n = case["n"]
        lo, hi = rank * n // world, (rank + 1) * n // world
        # ---- (a)+(b): one sweep ----
        sw = AllReduceSweep(case["x"][:, lo:hi], O.NIW, seed=99, global_offset=lo)
        set_params(sw, case)
        sw.sample_labels(False)
        sw.sample_sublabels()
        stats = sw.suff_stats()
        # ---- (c): SPMD fit ----
        def factory(x, kind, seed, goff):
            return AllReduceSweep(x, kind, seed=seed, global_offset=goff)
        out = H.fit(case["x"][:, lo:hi], 10.0, iters=30, seed=5, burnout=5, sweep_factory=factory, shard=(lo, n))
        smart = H.fit(case["x"][:, lo:hi], 10.0, iters=30, seed=5, burnout=5, sweep_factory=factory, shard=(lo, n),
                      smart_splits=True)
        q.put((rank, lo, hi, sw.get_labels(), sw.get_sublabels(), stats, out[0], len(out[1]), out[6], smart[6]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_process_sharded_sweep_and_fit():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=500) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case = make_niw_case(4, 5, 3000, seed=8)
    ref = O.OracleSweep(case["x"], O.NIW, seed=99)
    set_params(ref, case)
    ref.sample_labels(False)
    ref.sample_sublabels()
    rc, rsx, rsxx = ref.suff_stats()
    assert res[0][9] == res[1][9] and len(res[0][9]) == 30                  # smart splits: identical decisions on every rank
    for rank, lo, hi, lab, sub, stats, *_ in res:
        np.testing.assert_array_equal(lab, ref.get_labels()[lo:hi])        # (a) shard-invariant draws
        np.testing.assert_array_equal(sub, ref.get_sublabels()[lo:hi])
        np.testing.assert_array_equal(stats[0], rc)                         # (b) all-reduced statistics
        np.testing.assert_allclose(stats[1], rsx, rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(stats[2], rsxx, rtol=1e-12, atol=1e-9)
    single = H.fit(case["x"], 10.0, iters=30, seed=5, burnout=5,
                   sweep_factory=lambda x, kind, seed, goff: O.OracleSweep(x, kind, seed=seed, global_offset=goff))
    assert res[0][7] == res[1][7] == len(single[1])                         # (c) same K on every rank
    assert res[0][8] == res[1][8] == single[6]                              # same cluster-count history
    np.testing.assert_array_equal(np.concatenate([res[0][6], res[1][6]]), single[0])
