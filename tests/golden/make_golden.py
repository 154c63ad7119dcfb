#!/usr/bin/env python
"""Generate tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref/libsocp_ref.so,
built by `make -C oracle ref` in the container that has /root/reference).

    python tests/golden/make_golden.py

Every number is stored as float.hex() so the fixtures are bit-exact.  The scenarios restate the
reference's demo programs (tests/test*.cpp) -- see tests/scenarios.py for file:line citations.
The reference solves through oracle/minpack.c (clean-room MINPACK, pinned against scipy's MINPACK
in tests/test_oracle_minpack.py) because cminpack is not vendored with the reference.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import scenarios as S          # noqa: E402
from backends import RefBackend  # noqa: E402
from oracle import pyref as R   # noqa: E402


def hx(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 0:
        return float(a).hex()
    return [hx(v) for v in a]


def spec_hex(spec):
    out = dict(spec)
    for k in ("mparams", "time", "Xb", "x0"):
        out[k] = hx(spec[k])
    out["xtol"] = float(spec["xtol"]).hex()
    return out


def main():
    ref = RefBackend()
    rng = np.random.default_rng(20260000)
    G = dict(traj=[], points=[], residual=[], solve=[], cont_param=[], cont_boundary=[])

    def add_traj(model, mparams, t0, X0, tf, steps=None, sw=None):
        Xf = ref.traj(model, mparams, t0, X0, tf, steps, sw)
        G["traj"].append(dict(model=model, mparams=hx(mparams), steps=steps or S.STEPS[model],
                              t0=float(t0).hex(), tf=float(tf).hex(), X0=hx(X0), Xf=hx(Xf),
                              sw=hx(sw) if sw is not None else None))

    def add_point(model, mparams, t, X, sw=None):
        m = ref.model(model, mparams)
        if sw is not None:
            m.switching_times(sw)
        G["points"].append(dict(model=model, mparams=hx(mparams), t=float(t).hex(), X=hx(X),
                                sw=hx(sw) if sw is not None else None,
                                rhs=hx(m.rhs(t, X)), control=hx(m.control(t, X)),
                                H=hx(m.hamiltonian(t, X)[0])))

    # ---------------- trajectories and point evaluations ----------------
    gp = list(S.DEFAULTS[S.GODDARD]); gp[6] = 1.0
    gp0 = list(gp); gp0[2] = 0.0
    add_traj(S.GODDARD, gp, 0.0, S.GODDARD_XI, 0.1)          # SURVEY 8c known answer (KD=310)
    add_traj(S.GODDARD, gp0, 0.0, S.GODDARD_XI, 0.1)         # SURVEY 8c known answer (KD=0)
    Xi_b, _ = S.goddard_batch_inputs(6)
    for k in range(6):
        add_traj(S.GODDARD, gp0 if k % 2 else gp, 0.0, Xi_b[k], 0.03 + 0.01 * k)
    add_traj(S.GODDARD, gp, 0.1, S.GODDARD_XI, 0.1)          # tf == t0: zero steps (odeTools.cpp:135)
    add_traj(S.GODDARD, gp, 0.1, S.GODDARD_XI, 0.05)         # tf < t0: zero steps
    gs = list(S.DEFAULTS[S.GODDARD]); gs[6] = 0.0             # bang-singular-off structure
    Xs = S.GODDARD_XI.copy(); Xs[3:6] = [0.05, 0.01, 0.02]
    add_traj(S.GODDARD, gs, 0.0, Xs, 0.12, sw=[0.0227, 0.08])
    for t in (0.01, 0.05, 0.1):
        add_point(S.GODDARD, gs, t, Xs, sw=[0.0227, 0.08])
    add_point(S.GODDARD, gp, 0.02, S.GODDARD_XI)
    gsat = list(gp); gsat[6] = 0.01                            # saturated quadratic control
    add_point(S.GODDARD, gsat, 0.02, S.GODDARD_XI)

    cp = list(S.DEFAULTS[S.COVID19]); cp[0:3] = [3.4, 14.0, 5.0]
    add_traj(S.COVID19, cp, 0.0, S.COVID_XI, 1.5)             # SURVEY 8c known answer
    for k in range(4):
        X = np.r_[rng.dirichlet([5, 1, 1, 2]), rng.normal(size=4)]
        cpk = list(cp); cpk[4] = 0.05 + 0.1 * k; cpk[5] = 10.0 ** (k - 1)
        add_traj(S.COVID19, cpk, 0.0, X, 18.25)
        add_point(S.COVID19, cpk, 0.0, X)

    dp = list(S.DEFAULTS[S.DI])
    add_traj(S.DI, dp, 0.0, [0, 0, 0, 0, 0, 0, .01, .01, .01, .01, .01, .01], 10.0)
    for k in range(3):
        X = rng.normal(size=12) * (0.3 if k == 0 else 3.0)    # unsaturated / saturated control
        add_traj(S.DI, dp, 0.0, X, 10.0)
        add_point(S.DI, dp, 0.0, X)

    vp = list(S.DEFAULTS[S.VTOL]); vp[6] = 0.05
    Xv = np.array([20, 8, 5, 0.3, 0.2, 0.1, -0.03, 0.013, -0.003, -0.29, 0.1, -0.03])
    add_traj(S.VTOL, vp, 0.0, Xv, 4.0)
    for pos in ([10.0, 40.0, 18.0], [35.0, 20.0, 20.0], [48.0, 73.0, 25.5], [60.0, 30.0, 12.0]):
        X = Xv.copy(); X[:3] = pos
        add_traj(S.VTOL, vp, 0.0, X, 3.0)
        add_point(S.VTOL, vp, 0.0, X)
    vp2 = list(vp); vp2[0] = 1.0; vp2[11] = 0.5                # saturating u_max, sharper obstacles
    add_point(S.VTOL, vp2, 0.0, Xv)

    ip = list(S.DEFAULTS[S.INTERCEPTOR])
    Xi0 = np.array(S.INTERCEPTOR_INIT_XI + [0.01, -1, 0.5, 0.2, 100., 50.])
    add_traj(S.INTERCEPTOR, ip, 0.0, Xi0, 10.0)
    add_traj(S.INTERCEPTOR, ip, 0.0, Xi0, 30.0)               # two stages (t1 = 20 s)
    add_traj(S.INTERCEPTOR, ip, 25.0, Xi0, 40.0)              # unpowered only
    Xs = Xi0.copy(); Xs[2] = 1.5                              # near-vertical: chart 1 -> 2
    add_traj(S.INTERCEPTOR, ip, 0.0, Xs, 10.0)
    Xs = Xi0.copy(); Xs[2] = -1.52
    add_traj(S.INTERCEPTOR, ip, 0.0, Xs, 6.0)
    ip0 = list(ip); ip0[13] = 0.0
    add_traj(S.INTERCEPTOR, ip0, 0.0, Xi0, 10.0)
    add_point(S.INTERCEPTOR, ip, 3.0, Xi0)
    add_point(S.INTERCEPTOR, ip0, 3.0, Xi0)

    # ---------------- residuals and solves ----------------
    def add_residual(spec, x=None):
        f = ref.residual(spec, x)
        G["residual"].append(dict(spec=spec_hex(spec), x=hx(spec["x0"] if x is None else x), fvec=hx(f)))

    def add_solve(spec):
        r = ref.solve(spec)
        G["solve"].append(dict(spec=spec_hex(spec), x=hx(r["x"]), info=r["info"], nfev=r["nfev"]))
        print("solve %-22s info=%d nfev=%d" % (spec["name"], r["info"], r["nfev"]))
        return r

    def add_cont_param(spec, step, pname, goal):
        r = ref.continuation_param(spec, step, pname, goal)
        G["cont_param"].append(dict(spec=spec_hex(spec), step=step, pname=pname, goal=float(goal).hex(),
                                    x=hx(r["x"]), info=r["info"], solver_calls=r["solver_calls"],
                                    nfev_total=r["nfev_total"]))
        print("cont  %-22s %s->%g info=%d calls=%d nfev=%d" % (spec["name"], pname, goal, r["info"],
                                                             r["solver_calls"], r["nfev_total"]))
        return r

    def add_cont_boundary(spec, step, timed, Xd):
        r = ref.continuation_boundary(spec, step, timed, Xd)
        G["cont_boundary"].append(dict(spec=spec_hex(spec), step=step, timed=hx(timed), Xd=hx(Xd),
                                       x=hx(r["x"]), info=r["info"], solver_calls=r["solver_calls"],
                                       nfev_total=r["nfev_total"]))
        print("contB %-22s info=%d calls=%d nfev=%d" % (spec["name"], r["info"], r["solver_calls"],
                                                      r["nfev_total"]))
        return r

    def restart(spec, x, name, **changes):
        s = dict(spec)
        s["x0"] = [float(v) for v in x]
        s["name"] = name
        s.update(changes)
        return s

    # double integrator (tests/testDoubleIntegrator.cpp with modelOrder = 0 -> hybrd)
    di = S.di_problem()
    add_residual(di)
    r = add_solve(di)
    Xd = [list(di["Xb"][0]), list(di["Xb"][1])]
    Xd[1][1] = 20.0
    di2 = restart(di, r["x"], "di_cont_Xf")
    r2 = add_cont_boundary(di2, 1.0, di["time"], Xd)
    di3 = restart(di, r2["x"], "di_cont_muT", Xb=Xd)
    add_cont_param(di3, 1.0, "muT", 0.02)

    # Goddard (tests/testGoddard.cpp), 4 stages
    g1 = S.goddard_problem(lambda mp, a, b, c: ref.traj(S.GODDARD, mp, a, b, c, 10))
    add_residual(g1)
    r1 = add_solve(g1)
    g2 = restart(g1, r1["x"], "goddard_stage2_KD")
    r2 = add_cont_param(g2, 1.0, "KD", 310.0)
    mp2 = list(g1["mparams"]); mp2[2] = 310.0
    g3 = restart(g1, r2["x"], "goddard_stage3_mu2", mparams=mp2)
    r3 = add_cont_param(g3, 1.0, "mu2", 0.2)
    mp3 = list(mp2); mp3[6] = 0.2
    # re-mesh on the switching times (tests/testGoddard.cpp:115-145), through the reference itself
    m, s = ref.shooting(restart(g1, r3["x"], "tmp", mparams=mp3))
    vt, _ = s.solution()
    tf = vt[6]
    s1, s2 = 0.0227, 0.08
    vt = np.array([0.0, s1 / 2, s1, (s2 + s1) / 2, s2, (s2 + tf) / 2, tf])
    vX = np.array([s.move(t) for t in vt])
    mode_t = [S.FIXED, S.CONTINUOUS, S.FREE, S.CONTINUOUS, S.FREE, S.CONTINUOUS, S.FREE]
    mode_X = [[S.FIXED] * 7] + [[S.CONTINUOUS] * 7] * 5 + [S.GODDARD_MODE_XF]
    x4 = np.r_[vX[:6].reshape(-1), vt[2], vt[4], vt[6]]
    mp4 = list(mp3); mp4[6] = 0.0; mp4[7] = -1.0
    Xb4 = np.zeros((7, 7)); Xb4[:] = vX[:, :7]
    g4 = S.make_spec(S.GODDARD, 6, mode_t, mode_X, vt, Xb4, x4, 1e-6, mparams=mp4, name="goddard_stage4_singular")
    add_residual(g4)
    add_solve(g4)
    # a few perturbed batch members (config C2), single shooting and M=6
    Xi_b, xf_b = S.goddard_batch_inputs(3)
    for k in range(3):
        gk = S.goddard_problem(lambda mp, a, b, c: ref.traj(S.GODDARD, mp, a, b, c, 10), Xi=Xi_b[k], xf0=xf_b[k])
        gk["name"] = "goddard_batch_%d" % k
        add_solve(gk)

    # covid19 (tests/testCovid19.cpp): first solve, and the first boundary continuation
    c1 = S.covid_problem(lambda mp, a, b, c: ref.traj(S.COVID19, mp, a, b, c))
    add_residual(c1)
    rc = add_solve(c1)
    Xd = [list(r) for r in c1["Xb"]]
    Xd[20][3] = 0.7
    add_cont_boundary(restart(c1, rc["x"], "covid_cont_Rf"), 0.1, c1["time"], Xd)

    # interceptor (tests/testInterceptor.cpp): init problem, mu_gft continuation, scenario S3
    mi = ref.model(S.INTERCEPTOR, S.DEFAULTS[S.INTERCEPTOR][:13] + [0.0] + S.DEFAULTS[S.INTERCEPTOR][14:])
    Xa, _ = mi.init_analytical(0.0, np.r_[S.INTERCEPTOR_INIT_XI, np.zeros(6)], 10.0,
                               np.r_[S.INTERCEPTOR_INIT_XF, np.zeros(6)])
    G["interceptor_costate_guess"] = hx(Xa[6:])
    i1 = S.interceptor_init_problem(Xa[6:])
    add_residual(i1)
    ri = add_solve(i1)
    i2 = restart(i1, ri["x"], "interceptor_mu_gft")
    ri2 = add_cont_param(i2, 0.1, "mu_gft", 1.0)
    mpi = list(i1["mparams"]); mpi[13] = 1.0
    # boundary continuation to S3 / S2: previous data = init problem with the solved tf
    for sc in ("S3", "S2", "S1"):
        Xi_s, Xf_s = S.INTERCEPTOR_SCENARIOS[sc]
        m, s = ref.shooting(restart(i1, ri2["x"], "tmp", mparams=mpi))
        vt, vX = s.solution()                                     # GetSolution (testInterceptor.cpp:194-199)
        x0 = np.r_[vX[0], vt[1]]
        spec = S.make_spec(S.INTERCEPTOR, 1, i1["mode_t"], i1["mode_X"], vt, [vX[0][:6], vX[1][:6]], x0,
                           1e-8, mparams=mpi, name="interceptor_" + sc)
        add_cont_boundary(spec, 0.1, [0.0, 20.0], [Xi_s, Xf_s])

    # vtolUAV first leg (tests/testVtolUAV.cpp:163-232) + drag continuation
    v1 = S.vtol_first_problem()
    add_residual(v1)
    rv = add_solve(v1)
    add_cont_param(restart(v1, rv["x"], "vtol_cont_ca"), 0.1, "ca", 0.05)

    out = os.path.join(HERE, "golden.json")
    with open(out, "w") as f:
        json.dump(G, f, indent=0, separators=(",", ":"))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    if not R.available():
        sys.exit("oracle/_ref/libsocp_ref.so missing: run `make -C oracle ref` where /root/reference exists")
    main()
