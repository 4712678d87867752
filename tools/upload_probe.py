"""Development aid: where the time of GpuSweep(x) goes (C2: 128 MB of X)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dpmm_pkg
pkg = dpmm_pkg.load()
x = np.asfortranarray(np.random.default_rng(0).standard_normal((32, 1_000_000)).astype(np.float32))
torch.cuda.init(); torch.zeros(1, device="cuda"); torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); g = pkg.GpuSweep(x, pkg.NIW, seed=1); g.sync(); t1 = time.perf_counter()
    g.close(); t2 = time.perf_counter()
    xt = torch.from_numpy(x.T)   # [n, 32] C-contiguous view
    t3 = time.perf_counter(); d = xt.cuda(); torch.cuda.synchronize(); t4 = time.perf_counter()
    xp = xt.pin_memory(); t5 = time.perf_counter(); d2 = xp.cuda(non_blocking=True); torch.cuda.synchronize(); t6 = time.perf_counter()
    print(f"rep {rep}: GpuSweep create+upload {1e3*(t1-t0):.1f} ms, close {1e3*(t2-t1):.1f} ms | torch pageable H2D {1e3*(t4-t3):.1f} ms | pin {1e3*(t5-t4):.1f} ms, pinned H2D {1e3*(t6-t5):.1f} ms", flush=True)
