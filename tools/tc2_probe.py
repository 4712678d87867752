"""Per-kernel timing + candidate / overflow statistics of the D = 32 / 64 tensor-core label path on a bench case."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, dpmm_pkg
pkg = dpmm_pkg.load()
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
case = bench.build_case(name, 0, 0)
os.environ["DPMM_TC_STATS"] = "1"
g = pkg.GpuSweep(case["x"], case["kind"], seed=1)
bench.set_params(g, case)
g.set_labels(case["gt"]) if os.environ.get("PROBE_GT") else None
for _ in range(3):
    g.sample_labels(False); g.sample_sublabels(); g.suff_stats(fetch=False)
g.timing_enable(True)
for _ in range(steps):
    g.sample_labels(False); g.sample_sublabels(); g.suff_stats(fetch=False)
tim = g.timing_read()
out = {k: round(ms / steps * 1e3, 1) for k, (ms, c) in tim.items() if c}
out["tc_stats(points, exact evals, overflow)"] = g.tc_stats(overflow=True)
out["case"] = name; out["K"] = case["K"]
print(json.dumps(out))
