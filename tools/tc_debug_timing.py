"""Development aid: time the tensor-core label kernel with parts of its epilogue switched off (DPMM_TC_DEBUG)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dpmm_pkg
from tests.util import make_niw_case, set_params
pkg = dpmm_pkg.load()
case = make_niw_case(32, 20, 1_000_000, 1, spread=56)
g = pkg.GpuSweep(case["x"], case["kind"], seed=1)
set_params(g, case)
for mode in [0, 3, 4, 7]:
    os.environ["DPMM_TC_DEBUG"] = str(mode)
    for _ in range(2): g.sample_labels()
    g.sync(); g.timing_enable(True)
    for _ in range(10): g.sample_labels()
    g.sync(); t = g.timing_read(); g.timing_enable(False)
    print(f"dbg={mode}  label {t['label'][0] / max(t['label'][1], 1) * 1e3:7.1f} us", flush=True)
