"""Development aid: confirm the tensor-core label path ran and report its candidate statistics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DPMM_TC_STATS"] = "1"
import numpy as np
import dpmm_pkg
from tests.util import make_niw_case, set_params
pkg = dpmm_pkg.load()
sys.argv = [sys.argv[0]]
import bench
for name, case in (("random-order overlapping (tests.util)", make_niw_case(32, 20, 1_000_000, 1)), ("bench C2 (blocked order)", bench.build_case("c2"))):
    g = pkg.GpuSweep(case["x"], case["kind"], seed=1)
    set_params(g, case)
    g.sample_labels(False)
    npts, ncand = g.tc_stats()
    lab = g.get_labels()
    os.environ["DPMM_LABEL_TC"] = "0"
    g2 = pkg.GpuSweep(case["x"], case["kind"], seed=1)
    set_params(g2, case)
    g2.sample_labels(False)
    lab2 = g2.get_labels()
    os.environ["DPMM_LABEL_TC"] = "1"
    print(f"{name}: points {npts}, exact evaluations {ncand} ({ncand/max(npts,1):.2f} per point of K={case['K']}); "
          f"labels differing from the FFMA path: {(lab != lab2).sum()}")
    g.close(); g2.close()
