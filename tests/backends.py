"""Run a scenario spec (tests/scenarios.py) through a checker: the unmodified reference build
(oracle/_ref) or the plain-C restatement (oracle/liboracle.so).  Test infrastructure only."""
import os

import numpy as np

from oracle import pyoracle as O
from oracle import pyref as R
import scenarios as S

_OBS = None


def oracle_obstacles():
    global _OBS
    if _OBS is None:
        o = S.VTOL_OBSTACLES
        _OBS = O.make_obstacles(o["type"], o["pos"], o["rad"])
    return _OBS


def _obstacle_files(tmpdir=None):
    """Write the obstacle table in the reference's text format for the _ref build."""
    import tempfile
    d = tmpdir or tempfile.mkdtemp(prefix="socp_obs_")
    o = S.VTOL_OBSTACLES
    path = os.path.join(d, "obstacles")
    with open(path, "w") as f:
        f.write("n:\n\t%d\ntype:\n" % len(o["type"]))
        for t in o["type"]:
            f.write("\t%d\n" % int(t))
        f.write("position:\n")
        for p in o["pos"]:
            f.write("\t%r\t%r\t%r\n" % tuple(float(v) for v in p))
        f.write("radius:\n")
        for p in o["rad"]:
            f.write("\t%r\t%r\t%r\n" % tuple(float(v) for v in p))
    wp = os.path.join(d, "waypoints")
    with open(wp, "w") as f:
        f.write("n_wp:\n%d\nposition_wp:\n" % len(S.VTOL_WAYPOINTS))
        for p in S.VTOL_WAYPOINTS:
            f.write("%d\t%d\t%d\n" % tuple(p))
    return path, wp


class OracleBackend:
    name = "oracle"

    def problem(self, spec):
        model = spec["model"]
        p = O.OracleProblem(model, spec["M"], step_nbr=spec["steps"],
                            obstacles=oracle_obstacles() if model == S.VTOL else None)
        for k, v in enumerate(spec["mparams"]):
            p.set_param(k, v)
        p.set_mode(spec["mode_t"], spec["mode_X"])
        p.set_boundary(spec["time"], spec["Xb"])
        return p

    def traj(self, model, mparams, t0, X0, tf, steps=None, sw=None, noise_ulps=0.0, seed=1):
        p = O.OracleProblem(model, 1, step_nbr=steps,
                            obstacles=oracle_obstacles() if model == S.VTOL else None)
        p.p.noise_ulps = noise_ulps
        p.p.noise_state = seed
        for k, v in enumerate(mparams):
            p.set_param(k, v)
        if sw is not None:
            for k, v in enumerate(sw):
                p.p.sw[k] = v
            p.p.nsw = len(sw)
        return p.traj(t0, X0, tf)

    def residual(self, spec, x=None):
        return self.problem(spec).residual(spec["x0"] if x is None else x)

    def fdjac(self, spec, x=None, noise_ulps=0.0, seed=1):
        p = self.problem(spec)
        p.p.noise_ulps = noise_ulps
        p.p.noise_state = seed
        return p.fdjac(spec["x0"] if x is None else x)

    def solve(self, spec, maxfev=10000):
        return self.problem(spec).solve(spec["x0"], xtol=spec["xtol"], maxfev=maxfev)

    def continuation_param(self, spec, step, pname, goal):
        p = self.problem(spec)
        r = p.continuation_param(spec["x0"], step, S.pidx(spec["model"], pname), goal, xtol=spec["xtol"])
        r["mparams"] = p.params(len(spec["mparams"]))
        return r

    def continuation_boundary(self, spec, step, timed, Xd):
        p = self.problem(spec)
        return p.continuation_boundary(spec["x0"], step, spec["time"], spec["Xb"], timed, Xd,
                                       xtol=spec["xtol"])


class RefBackend:
    name = "reference"

    def __init__(self):
        self._files = None

    def model(self, model, mparams, steps=None):
        kw = {}
        if model == S.VTOL:
            if self._files is None:
                self._files = _obstacle_files()
            kw = dict(obstacle_file=self._files[0], wp_file=self._files[1])
        m = R.RefModel(model, model_order=0, step_nbr=steps or 0, **kw)
        for name, v in zip(S.PARAMS[model], mparams):
            if model == S.INTERCEPTOR and name in ("r_2p", "t_2p"):
                continue
            m.set(name, v)
        return m

    def shooting(self, spec):
        model, M, n = spec["model"], spec["M"], S.DIM[spec["model"]]
        if spec["steps"] != S.STEPS[model] and model != S.GODDARD:
            raise ValueError("the reference hard-codes stepNbr for this model")
        m = self.model(model, spec["mparams"], spec["steps"])
        s = R.RefShooting(m, M, 1)
        s.set_precision(spec["xtol"])
        s.set_mode(spec["mode_t"], spec["mode_X"])
        x0 = np.asarray(spec["x0"], dtype=np.float64)
        # InitShooting(vt, vX): tab_param <- node (state, costate) blocks and FREE times
        vt = np.array(spec["time"], dtype=np.float64)
        k = 2 * n * M
        for j in range(M + 1):
            if spec["mode_t"][j] == S.FREE:
                vt[j] = x0[k]
                k += 1
        vX = np.zeros((M + 1, 2 * n))
        for j in range(M):
            vX[j] = x0[2 * n * j:2 * n * (j + 1)]
        vX[M, :n] = spec["Xb"][M]
        s.init_v(vt, vX)
        # boundary data actually used by the residual: desired state -> data->X on SolveOCP(0)
        Xd = np.zeros((M + 1, 2 * n))
        Xd[:, :n] = np.asarray(spec["Xb"], dtype=np.float64)
        s.desired_v(np.array(spec["time"], dtype=np.float64), Xd)
        assert np.array_equal(s.params(), x0)
        return m, s

    def traj(self, model, mparams, t0, X0, tf, steps=None, sw=None):
        m = self.model(model, mparams, steps)
        if sw is not None:
            m.switching_times(sw)
        return m.traj(t0, X0, tf)

    def residual(self, spec, x=None):
        _, s = self.shooting(spec)
        return s.residual(spec["x0"] if x is None else x)

    def solve(self, spec):
        _, s = self.shooting(spec)
        info = s.solve(0.0)
        nfev, _ = s.call_number()
        # on failure tab_param keeps the guess (shooting.cpp:588); report it as the reference does
        return dict(x=s.params(), info=info, nfev=nfev)

    def continuation_param(self, spec, step, pname, goal):
        m, s = self.shooting(spec)
        R.log_clear()
        info = s.solve_param(step, pname, goal)
        log = R.log()
        return dict(x=s.params(), info=info, solver_calls=len(log),
                    nfev_total=sum(c[1] for c in log), mparams=m.params())

    def continuation_boundary(self, spec, step, timed, Xd):
        """spec['time'] / spec['Xb'] are the previous boundary data (time_prec, X_prec)."""
        m, s = self.shooting(spec)
        n = S.DIM[spec["model"]]
        XD = np.zeros((spec["M"] + 1, 2 * n))
        XD[:, :n] = np.asarray(Xd, dtype=np.float64)
        s.desired_v(np.asarray(timed, dtype=np.float64), XD)
        R.log_clear()
        info = s.solve(step)
        log = R.log()
        return dict(x=s.params(), info=info, solver_calls=len(log), nfev_total=sum(c[1] for c in log))
