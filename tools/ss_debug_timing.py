"""Development aid: time the fused kernel with parts switched off (DPMM_SS_DEBUG bit mask).
Needs a library built with DPMM_BUILD_DEBUG_SWITCHES=1 python __graft_entry__.py (force a rebuild: touch csrc/*.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dpmm_pkg
from tests.util import make_niw_case, set_params
pkg = dpmm_pkg.load()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
case = make_niw_case(32, 20, n, 1, spread=56)
g = pkg.GpuSweep(case["x"], case["kind"], seed=1)
set_params(g, case)
g.sample_labels()
for mode in [0, 1, 2, 4, 8, 16, 32, 128, 256, 512, 1024 + 2, 15, 63]:
    os.environ["DPMM_SS_DEBUG"] = str(mode)
    for _ in range(2): g.sample_sublabels()
    g.sync(); g.timing_enable(True)
    for _ in range(10): g.sample_sublabels()
    g.sync(); t = g.timing_read(); g.timing_enable(False)
    print(f"dbg={mode:3d}  sublabel {t['sublabel'][0] / max(t['sublabel'][1], 1) * 1e3:7.1f} us", flush=True)
