"""A/B check of the two builds of the Broyden phase on the same batch: SOCP_BROYDEN=fused (one kernel, Q swept
twice) against the default split build (Q pass + chain kernel).  They are meant to agree bit for bit."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import socp_b200 as sb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = sb.Engine(0)
out = {}
for wl in (bench.wl_goddard, bench.wl_goddard_warm):
    w = wl(eng, B, 20260002)
    for mode in ("fused", "split"):
        os.environ["SOCP_BROYDEN"] = mode
        x = np.ascontiguousarray(w.x0).copy()
        t0 = time.perf_counter()
        r = eng.solve_batch(w.shape, w.mp, w.time, w.Xb, x, xtol=w.xtol, maxfev=10000)
        dt = time.perf_counter() - t0
        out[(w.name, mode)] = (r["x"].copy(), r["info"].copy(), r["nfev"].copy(), r["fnorm"].copy())
        print("%-13s %-5s %.3f s  converged %.4f  mean nfev %.1f" % (w.name, mode, dt, (r["info"] == 1).mean(), r["nfev"].mean()))
    a, b = out[(w.name, "fused")], out[(w.name, "split")]
    same = [np.array_equal(u, v) for u, v in zip(a, b)]
    print("%-13s fused vs split: x %s info %s nfev %s fnorm %s; differing problems: %d" %
          (w.name, *same, int(np.sum((a[1] != b[1]) | (a[2] != b[2])))))
