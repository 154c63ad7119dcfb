"""Per-phase cycle counters of the Powell-hybrid kernels (SOCP_PHASE_CLOCKS=1) and, with SOCP_ROUND_LOG=<file>, the
per-round kernel times of one batched Goddard solve.  usage: phase_clocks.py [batch] [workload]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import socp_b200 as sb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
wl = bench.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "goddard"][0]
eng = sb.Engine(0)
w = wl(eng, B, 20260002)
x = np.ascontiguousarray(w.x0).copy()
eng.solve_batch(w.shape, w.mp, w.time, w.Xb, x, xtol=w.xtol)          # warm-up
eng.reset_stats()
eng.set_profiling(True)
x = np.ascontiguousarray(w.x0).copy()
r = eng.solve_batch(w.shape, w.mp, w.time, w.Xb, x, xtol=w.xtol)
st = eng.stats()
print({k: round(v, 1) for k, v in st.items()})
print("converged %.4f mean nfev %.1f" % ((r["info"] == 1).mean(), r["nfev"].mean()))
