"""A/B check of the Jacobian-phase routines on the same batch: SOCP_JAC=old (one barrier-separated reflector at a
time in shared memory) against the default (register-window panels).  They are meant to agree bit for bit."""
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
for wl in (bench.wl_goddard_warm, bench.wl_goddard):
    w = wl(eng, B, 20260002)
    for mode in ("old", "window"):
        os.environ["SOCP_JAC"] = mode
        x = np.ascontiguousarray(w.x0).copy()
        t0 = time.perf_counter()
        r = eng.solve_batch(w.shape, w.mp, w.time, w.Xb, x, xtol=w.xtol, maxfev=10000)
        dt = time.perf_counter() - t0
        out[(w.name, mode)] = (r["x"].copy(), r["info"].copy(), r["nfev"].copy(), r["fnorm"].copy())
        print("%-13s %-6s %.3f s  converged %.4f  mean nfev %.1f" % (w.name, mode, dt, (r["info"] == 1).mean(), r["nfev"].mean()), flush=True)
    a, b = out[(w.name, "old")], out[(w.name, "window")]
    same = [bool(np.array_equal(u, v)) for u, v in zip(a, b)]
    rel = np.max(np.abs(a[0] - b[0]) / np.maximum(np.abs(a[0]), 1e-300))
    print("%-13s old vs window: x %s info %s nfev %s fnorm %s; differing problems: %d; max rel dx %.2e" %
          (w.name, *same, int(np.sum((a[1] != b[1]) | (a[2] != b[2]))), rel), flush=True)
