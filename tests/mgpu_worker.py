"""torchrun worker for tests/test_gpu_multi.py: each rank owns one contiguous shard on its own GPU;
rank 0 additionally runs the whole data set on one GPU and compares."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dpmm_pkg  # noqa: E402
from tests.util import make_niw_case, set_params  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = dpmm_pkg.load()
case = make_niw_case(32, 8, 200000, seed=3)
n = case["n"]
lo, hi = rank * n // world, (rank + 1) * n // world
g = pkg.GpuSweep(case["x"][:, lo:hi], pkg.NIW, seed=42, global_offset=lo, device=local)
ids = [pkg.GpuSweep.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
g.comm_init(ids[0], rank, world)
set_params(g, case)
g.sample_labels(False)
g.sample_sublabels()
counts, sx, sxx = g.suff_stats()
lab, sub = g.get_labels(), g.get_sublabels()
# smart splits across shards (dpmm_smart_*): percentiles reduced with min / max, per-side sums summed over ranks
vdir = np.ones(32) / np.sqrt(32.0)
sm_lo, sm_hi, sm_cnt = g.smart_project(1, vdir, np.zeros(32))
sm_sums = g.smart_kmeans_iter(sm_lo, sm_hi)
gathered = [None] * world
dist.all_gather_object(gathered, (lo, hi, lab, sub))
if rank == 0:
    ref = pkg.GpuSweep(case["x"], pkg.NIW, seed=42, global_offset=0, device=local)
    set_params(ref, case)
    ref.sample_labels(False)
    ref.sample_sublabels()
    rc, rsx, rsxx = ref.suff_stats()
    full_lab = np.concatenate([t[2] for t in sorted(gathered)])
    full_sub = np.concatenate([t[3] for t in sorted(gathered)])
    np.testing.assert_array_equal(full_lab, ref.get_labels())
    np.testing.assert_array_equal(full_sub, ref.get_sublabels())
    np.testing.assert_array_equal(counts, rc)
    # runs are accumulated in Float32 (<= 1024 points) before the Float64 atomics, and the run
    # boundaries depend on the sharding: agreement is to Float32 rounding, not bit-exact
    np.testing.assert_allclose(sx, rsx, rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(sxx, rsxx, rtol=1e-5, atol=1e-5 * np.abs(rsxx).max())
    from oracle import dpmm_oracle as O
    sel = full_lab == 1
    t = vdir @ case["x"][:, sel].astype(np.float64)
    assert sm_cnt == int(sel.sum())
    shard_t = [vdir @ case["x"][:, a:b][:, full_lab[a:b] == 1].astype(np.float64) for a, b, *_ in sorted(gathered)]
    want_lo = min(O.julia_percentile(s_, 0.10) for s_ in shard_t if s_.size > 1)
    want_hi = max(O.julia_percentile(s_, 0.90) for s_ in shard_t if s_.size > 1)
    assert abs(sm_lo - want_lo) < 1e-9 and abs(sm_hi - want_hi) < 1e-9
    left = np.abs(t - sm_lo) < np.abs(t - sm_hi)
    np.testing.assert_allclose(sm_sums, [t[left].sum(), left.sum(), t[~left].sum(), (~left).sum()], rtol=1e-10)
    print("MGPU_OK", counts[:, 0].sum(), flush=True)
g.close()
dist.barrier()
dist.destroy_process_group()
