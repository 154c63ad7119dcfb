"""Scenario specs shared by every parity test (pure data; no oracle / product imports).

A *spec* is a plain dict describing one shooting problem exactly as the reference's
`shooting::data_struct` + model parameters hold it at the moment hybrd is called
(/root/reference/src/socp/shooting.cpp:21-54):

    model      int   0 goddard, 1 doubleIntegrator, 2 covid19, 3 vtolUAV, 4 interceptor
    M          int   numMulti
    steps      int   RK4 steps per segment (per stage for the interceptor)
    mode_t     [M+1]        FIXED/FREE/CONTINUOUS per node time
    mode_X     [M+1][dim]   per node state component
    mparams    [np]         model parameter block (index maps: PARAMS below)
    time       [M+1]        data->time  (fixed node times; FREE entries are ignored)
    Xb         [M+1][dim]   data->X[i][0..dim)  boundary / waypoint states
    x0         [P]          unknowns (tab_param)
    xtol       float

The demo scenarios restate /root/reference/tests/test*.cpp (file:line cited per builder).
"""
import numpy as np

GODDARD, DI, COVID19, VTOL, INTERCEPTOR = range(5)
FIXED, FREE, CONTINUOUS = 0, 1, 2
DIM = [7, 6, 4, 6, 6]
STEPS = [10, 30, 1000, 100, 50]
PARAMS = {
    GODDARD: ["C", "b", "KD", "kr", "u_max", "mu1", "mu2", "singularControl"],
    DI: ["u_max", "a_max", "muT"],
    COVID19: ["R0", "Tinf", "Tinc", "N", "Imax", "muI", "umin", "umax"],
    VTOL: ["u_max", "a_max", "alphaT", "alphaV", "invSigmaXwp", "Vd", "ca", "nWP_tot", "nWP",
           "phiObs", "psiWP", "muObs", "sigmaWP"],
    INTERCEPTOR: ["c0", "hr", "d0", "eta", "propellant_mass", "empty_mass", "q", "ve", "alpha_max",
                  "u_max", "a_max", "r_2p", "t_2p", "mu_gft", "muT", "muV", "muC"],
}
# model constructor defaults (goddard.hpp:29-36, doubleIntegrator.cpp:30-32, covid19.cpp:29-36,
# vtolUAV.cpp:27-35 + obstacle.cpp:45-48, interceptor.cpp:36-50; r_2p/t_2p are never read)
DEFAULTS = {
    GODDARD: [3.5, 7.0, 310.0, 500.0, 1.0, 1.0, 0.0, -1.0],
    DI: [1.0, 1.0, 0.01],
    COVID19: [4.0, 10.0, 5.0, 1.0, 0.1, 1.0, -10.0, 20.0],
    VTOL: [10.0, 0.3, 0.05, 0.0, 1.0 / 60, 1.0, 0.0, 0.0, 0.0, 1.0, 0.03, 1.0, 2.5],
    INTERCEPTOR: [0.00075, 7500.0, 0.00005, 0.442, 200.0, 200.0, 10.0, 1500.0, np.pi / 6, 1.0,
                  1500.0, 0.0, 0.0, 1.0, 0.0, 1.0, 0.0],
}
R_EARTH = 6378145.0

# data/vtolUAV/obstacles of the reference (9 boxes: type, position, radius)
VTOL_OBSTACLES = dict(
    type=[1.0] * 9,
    pos=[[7.5, 49.0, 20], [62.0, 49.0, 20], [35.0, 17.5, 20], [73.0, 17.5, 20], [39.5, 51.0, 20],
         [72.0, 51.0, 20], [24.5, 73.0, 20], [48.0, 73.0, 33], [48.0, 73.0, 7]],
    rad=[[4.0, 43.5, 20], [3.5, 43.5, 20], [23.5, 4.0, 20], [7.5, 4.0, 20], [19.0, 6.5, 20],
         [6.5, 6.5, 20], [13.0, 4.5, 20], [10.5, 4.5, 7], [10.5, 4.5, 7]],
)
# data/vtolUAV/waypoints (first entries; the path continuation adds one per stage)
VTOL_WAYPOINTS = [[20, 8, 5], [29, 5, 6], [39, 3, 7], [48, 2, 8], [57, 1, 9], [67, 1, 9], [76, 2, 9],
                  [81, 7, 9], [84, 15, 9], [84, 25, 10], [83, 34, 11], [82, 44, 12], [81, 53, 13],
                  [79, 62, 14], [76, 69, 15], [73, 78, 16], [70, 85, 17], [68, 93, 18], [62, 97, 19],
                  [55, 96, 20], [54, 87, 21], [52, 79, 22], [48, 72, 23], [43, 66, 24], [35, 63, 25],
                  [25, 63, 26], [17, 59, 27], [15, 50, 28], [17, 40, 29], [24, 37, 30], [32, 35, 31],
                  [40, 33, 32], [47, 30, 33]]


def pidx(model, name):
    return PARAMS[model].index(name)


def num_param(spec):
    return 2 * DIM[spec["model"]] * spec["M"] + sum(1 for m in spec["mode_t"] if m == FREE)


def default_modes(model, M, mode_tf, mode_Xf):
    """shooting::SetMode(mode_tf, mode_Xf) (shooting.cpp:165-182)."""
    n = DIM[model]
    mode_t = [FIXED] + [CONTINUOUS] * (M - 1) + [mode_tf]
    mode_X = [[FIXED] * n] + [[CONTINUOUS] * n for _ in range(M - 1)] + [list(mode_Xf)]
    return mode_t, mode_X


def make_spec(model, M, mode_t, mode_X, time, Xb, x0, xtol, mparams=None, steps=None, name=""):
    return dict(name=name, model=model, M=M, steps=steps or STEPS[model],
                mode_t=[int(v) for v in mode_t], mode_X=[[int(v) for v in r] for r in mode_X],
                mparams=[float(v) for v in (mparams if mparams is not None else DEFAULTS[model])],
                time=[float(v) for v in time], Xb=[[float(v) for v in r] for r in Xb],
                x0=[float(v) for v in x0], xtol=float(xtol))


def init_guess(traj, model, M, ti, Xi, tf, mode_t):
    """shooting::InitShooting(ti, Xi, tf, Xf) (shooting.cpp:202-245): node states by integrating
    from (ti, Xi) to each node time with `traj(t0, X0, tf) -> Xf`; FREE times appended."""
    n = DIM[model]
    time = [ti + i * (tf - ti) / M for i in range(M + 1)]
    x0 = list(Xi[:2 * n])
    for i in range(1, M):
        x0 += list(traj(ti, Xi, time[i]))
    for j in range(M + 1):
        if mode_t[j] == FREE:
            x0.append(time[j])
    return time, np.array(x0)


# ---- demo problems ---------------------------------------------------------------------------

def di_problem():
    """tests/testDoubleIntegrator.cpp:25-61 (modelOrder irrelevant for the residual itself)."""
    Xi = np.array([0, 0, 0, 0, 0, 0, .01, .01, .01, .01, .01, .01])
    Xf = np.zeros(6)
    Xf[0], Xf[1] = 10.0, 15.0
    mode_t, mode_X = default_modes(DI, 1, FREE, [FIXED] * 6)
    x0 = np.r_[Xi, 10.0]
    return make_spec(DI, 1, mode_t, mode_X, [0.0, 10.0], [Xi[:6], Xf], x0, 1e-8, name="di_free_tf")


GODDARD_XI = np.array([0.999949994, 0.0001, 0.01, 1e-10, 1e-10, 1e-10, 1.0] + [0.1] * 7)
GODDARD_MODE_XF = [FIXED, FIXED, FIXED, FREE, FREE, FREE, FREE]


def goddard_problem(traj, M=6, KD_init=310.0, mu2=1.0, KD=0.0, Xi=None, xf0=1.01):
    """tests/testGoddard.cpp:28-99: InitShooting runs with the constructor's KD=310 and mu2=1,
    KD is set to 0 afterwards.  `traj(mparams, t0, X0, tf)` integrates one segment."""
    Xi = GODDARD_XI if Xi is None else np.asarray(Xi, dtype=np.float64)
    Xf = np.zeros(7)
    Xf[0] = xf0
    mode_t, mode_X = default_modes(GODDARD, M, FREE, GODDARD_MODE_XF)
    mp_init = list(DEFAULTS[GODDARD])
    mp_init[pidx(GODDARD, "mu2")] = mu2
    mp_init[pidx(GODDARD, "KD")] = KD_init
    time, x0 = init_guess(lambda a, b, c: traj(mp_init, a, b, c), GODDARD, M, 0.0, Xi, 0.1, mode_t)
    mp = list(mp_init)
    mp[pidx(GODDARD, "KD")] = KD
    Xb = np.zeros((M + 1, 7))
    Xb[0] = Xi[:7]
    Xb[M] = Xf
    return make_spec(GODDARD, M, mode_t, mode_X, time, Xb, x0, 1e-6, mparams=mp, name="goddard_stage1")


COVID_XI = np.array([0.93, 0.003, 0.01, 0.057, -0.001, 0.001, 0.0, 0.0])


def covid_problem(traj, M=20, tf=30.0, Rf=0.6, steps=1000, mparams=None):
    """tests/testCovid19.cpp:42-78."""
    mp = list(DEFAULTS[COVID19]) if mparams is None else list(mparams)
    if mparams is None:
        mp[0], mp[1], mp[2] = 3.4, 14.0, 5.0
    mode_t, mode_X = default_modes(COVID19, M, FIXED, [FREE, FREE, FREE, FIXED])
    time, x0 = init_guess(lambda a, b, c: traj(mp, a, b, c), COVID19, M, 0.0, COVID_XI, tf, mode_t)
    Xb = np.zeros((M + 1, 4))
    Xb[0] = COVID_XI[:4]
    Xb[M][3] = Rf
    return make_spec(COVID19, M, mode_t, mode_X, time, Xb, x0, 1e-8, mparams=mp, steps=steps,
                     name="covid_stage1")


INTERCEPTOR_INIT_XI = [1000.0, 1000.0, np.pi / 4, 0.0, 5454661 / R_EARTH, 46086 / R_EARTH]
INTERCEPTOR_INIT_XF = [6000.0, 1000.0, 0.01 * np.pi, 0.01 * np.pi, (5454661 + 27829.0) / R_EARTH,
                       46086 / R_EARTH]
INTERCEPTOR_SCENARIOS = {   # tests/testInterceptor.cpp:36-97: (Xi, Xf), ti=0, tf=20
    "S1": ([3000.0, 1000.0, -np.pi / 6, 0.0, 5454661 / R_EARTH, 46086 / R_EARTH],
           [12000.0, 1000.0, 0.0, np.pi / 8, 5475000 / R_EARTH, 42000 / R_EARTH]),
    "S2": ([3000.0, 1000.0, np.pi / 4, 0.0, 5454661 / R_EARTH, 46086 / R_EARTH],
           [12000.0, 1000.0, -np.pi / 4, -np.pi / 2, 5485000 / R_EARTH, 36178 / R_EARTH]),
    "S3": ([3000.0, 1000.0, 0.0, 0.0, 5454661 / R_EARTH, 46086 / R_EARTH],
           [3000.0, 1000.0, 0.0, 0.0, 5485000 / R_EARTH, 46086 / R_EARTH]),
}


def interceptor_init_problem(costate_guess):
    """tests/testInterceptor.cpp:150-186: the initialisation problem with mu_gft = 0 and the
    analytic costate guess (interceptor::InitAnalytical, interceptor.cpp:844-955)."""
    mp = list(DEFAULTS[INTERCEPTOR])
    mp[pidx(INTERCEPTOR, "mu_gft")] = 0.0
    mode_t, mode_X = default_modes(INTERCEPTOR, 1, FREE, [FIXED, FREE, FIXED, FIXED, FIXED, FIXED])
    x0 = np.r_[INTERCEPTOR_INIT_XI, costate_guess, 10.0]
    return make_spec(INTERCEPTOR, 1, mode_t, mode_X, [0.0, 10.0],
                     [INTERCEPTOR_INIT_XI, INTERCEPTOR_INIT_XF], x0, 1e-8, mparams=mp,
                     name="interceptor_init")


def vtol_first_problem():
    """tests/testVtolUAV.cpp:163-214: first leg WP0 -> WP1, M=1, tf free, final velocity free."""
    wp = VTOL_WAYPOINTS
    mode_t = [FIXED, FREE]
    mode_X = [[FIXED] * 6, [FREE] * 6]
    mp = list(DEFAULTS[VTOL])
    mp[pidx(VTOL, "nWP_tot")] = len(wp) - 1
    mp[pidx(VTOL, "nWP")] = 0
    v0 = [0.001] * 3
    dX = np.array(wp[1], float) - np.array(wp[0], float)
    normDX = np.sqrt(dX[0] * dX[0] + dX[1] * dX[1] + dX[2] * dX[2])
    tf = pow(4.5 * normDX * normDX / mp[pidx(VTOL, "alphaT")], 0.25)
    p0 = [-3 * dX[0] / tf / tf / tf, -3 * dX[1] / tf / tf / tf, -3 * dX[2] / tf / tf / tf,
          -3 * dX[0] / tf / tf, -3 * dX[1] / tf / tf, -3 * dX[2] / tf / tf]
    x0 = np.r_[wp[0], v0, p0, tf]
    Xb = [list(wp[0]) + v0, list(wp[1]) + v0]
    return make_spec(VTOL, 1, mode_t, mode_X, [0.0, tf], Xb, x0, 1e-4, mparams=mp, name="vtol_wp1")


# ---- synthetic batches -------------------------------------------------

def goddard_batch_inputs(B, seed=20260002):
    """Config C2: perturbed Goddard problems.  Returns (Xi[B,14], xf0[B]) with
    Xi[0..2] +-1e-3 relative, mass +-1 %, target radius 1.01 +- 0.002; costate guess 0.1."""
    rng = np.random.default_rng(seed)
    Xi = np.tile(GODDARD_XI, (B, 1))
    Xi[:, 0:3] *= 1.0 + 1e-3 * rng.uniform(-1, 1, size=(B, 3))
    Xi[:, 6] *= 1.0 + 1e-2 * rng.uniform(-1, 1, size=B)
    xf0 = 1.01 + 0.002 * rng.uniform(-1, 1, size=B)
    return Xi, xf0
