"""Trajectory parity of a build of the library against the golden trajectories recorded from the unmodified
reference (tests/golden/golden.json): worst norm-relative and component-relative error per model.
Run once per build:   python tools/parity_builds.py            (default build, FMA contraction on)
                      SOCP_LIB=socp_b200/libsocp_b200_nofma.so python tools/parity_builds.py     (-fmad=false)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util as G  # noqa: E402
import scenarios as S  # noqa: E402
import socp_b200 as sb  # noqa: E402

eng = sb.Engine(0)
o = S.VTOL_OBSTACLES
eng.set_obstacles(o["type"], o["pos"], o["rad"])
worst = {}
for e in G.golden()["traj"]:
    model = e["model"]
    X0, want = G.unhex(e["X0"]), G.unhex(e["Xf"])
    sw = G.unhex(e["sw"]) if e.get("sw") else None
    got = eng.traj_batch(model, np.array(G.unhex(e["mparams"])), G.unhex(e["t0"]), X0[None, :], G.unhex(e["tf"]), e.get("steps", 0),
                         sw=None if sw is None else np.asarray(sw)[None, :])[0]
    scale = np.max(np.abs(want))
    nrel = np.max(np.abs(got - want)) / scale
    big = np.abs(want) >= 1e-6 * scale
    crel = np.max(np.abs(got - want)[big] / np.abs(want)[big])
    w = worst.setdefault(sb.MODEL_NAMES[model], [0.0, 0.0, 0])
    w[0], w[1], w[2] = max(w[0], nrel), max(w[1], crel), w[2] + 1
print("build:", os.environ.get("SOCP_LIB", "socp_b200/libsocp_b200.so (default, -fmad=true)"))
for name, (nrel, crel, cnt) in worst.items():
    print("  %-18s %2d golden trajectories: norm-relative %.2e, component-relative (|x_i| >= 1e-6 |x|) %.2e" % (name, cnt, nrel, crel))
