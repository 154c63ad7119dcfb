"""Golden vectors of the analytic-Jacobian path (modelOrder == 1), recorded from the UNMODIFIED reference
(oracle/_ref: src/socp/shooting.cpp + src/models/doubleIntegrator compiled where they lie).
Run in the authoring container:  python tests/golden/make_golden_hybrj.py  ->  tests/golden/golden_hybrj.json"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import scenarios as S                       # noqa: E402
from oracle import pyref                    # noqa: E402
from test_oracle_hybrj import di_wp_spec, ref_shooting   # noqa: E402


def hexv(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float64).reshape(-1)]


def main():
    assert pyref.available(), "needs oracle/_ref (make -C oracle ref)"
    out = {"traj_var": [], "jacobian": [], "solve": []}
    m = pyref.RefModel(S.DI, model_order=1, step_nbr=0)
    rng = np.random.default_rng(3)
    for _ in range(3):
        X = np.zeros(156)
        X[:12] = rng.uniform(-1, 1, 12)
        X[12:] = np.eye(12).reshape(-1) + 0.1 * rng.uniform(-1, 1, 144)
        tf = float(rng.uniform(1, 20))
        out["traj_var"].append(dict(X0=hexv(X), tf=tf.hex(), Xf=hexv(m.traj(0.0, X, tf, is_jac=1))))
    for name, make in (("di_free_tf", S.di_problem), ("di_wp", di_wp_spec)):
        spec = make()
        _, s = ref_shooting(spec)
        x = np.array(spec["x0"], dtype=np.float64)
        out["jacobian"].append(dict(name=name, x=hexv(x), J=hexv(s.jacobian(x))))
        info = s.solve(0.0)
        nfev, njev = s.call_number()
        out["solve"].append(dict(name=name, info=int(info), nfev=int(nfev), njev=int(njev), x=hexv(s.params())))
    with open(os.path.join(HERE, "golden_hybrj.json"), "w") as f:
        json.dump(out, f, indent=0)
    print({k: len(v) for k, v in out.items()}, [(e["name"], e["info"], e["nfev"], e["njev"]) for e in out["solve"]])


if __name__ == "__main__":
    main()
