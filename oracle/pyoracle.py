"""ctypes front-end to the plain-C restatement of the shooting path (oracle/socp_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's cpu_baseline / reference arm and
__graft_entry__.smoke(); never by the socp_b200 package.
"""
import ctypes
import os

import numpy as np

from . import pyminpack

GODDARD, DI, COVID19, VTOL, INTERCEPTOR = range(5)
FIXED, FREE, CONTINUOUS = 0, 1, 2
STATE_DIM = [7, 6, 4, 6, 6]
MAX_DIM, MAX_NODES, MAX_OBS = 7, 64, 32

_dp = ctypes.POINTER(ctypes.c_double)


class Obstacles(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int), ("type", ctypes.c_double * MAX_OBS),
                ("pos", (ctypes.c_double * 3) * MAX_OBS), ("rad", (ctypes.c_double * 3) * MAX_OBS)]


class Problem(ctypes.Structure):
    _fields_ = [("model_id", ctypes.c_int), ("dim", ctypes.c_int), ("num_multi", ctypes.c_int),
                ("step_nbr", ctypes.c_int),
                ("mode_t", ctypes.c_int * MAX_NODES),
                ("mode_X", (ctypes.c_int * MAX_DIM) * MAX_NODES),
                ("mparams", ctypes.c_double * 20),
                ("time", ctypes.c_double * MAX_NODES),
                ("Xb", (ctypes.c_double * MAX_DIM) * MAX_NODES),
                ("obs", ctypes.POINTER(Obstacles)),
                ("sw", ctypes.c_double * MAX_NODES), ("nsw", ctypes.c_int),
                ("chart", ctypes.c_int), ("stage", ctypes.c_int), ("rk4_steps", ctypes.c_long),
                ("noise_ulps", ctypes.c_double), ("noise_state", ctypes.c_ulonglong),
                ("integrator", ctypes.c_int), ("ode_tol", ctypes.c_double),
                ("dopri_steps", ctypes.c_long), ("dopri_rejected", ctypes.c_long)]


_SIGS = False


def lib():
    global _SIGS
    L = pyminpack.lib()
    if not _SIGS:
        pp = ctypes.POINTER(Problem)
        ci, cd = ctypes.c_int, ctypes.c_double
        L.so_problem_init.argtypes = [pp, ci, ci]
        L.so_num_param.argtypes = [pp]
        L.so_num_param.restype = ci
        L.so_default_steps.argtypes = [ci]
        L.so_default_steps.restype = ci
        L.so_rhs.argtypes = [pp, cd, _dp, _dp]
        L.so_control.argtypes = [pp, cd, _dp, _dp]
        L.so_control.restype = ci
        L.so_hamiltonian.argtypes = [pp, cd, _dp]
        L.so_hamiltonian.restype = cd
        L.so_obstacle_eval.argtypes = [pp, _dp, _dp, _dp]
        L.so_traj.argtypes = [pp, cd, _dp, cd, _dp]
        L.so_timeline.argtypes = [pp, _dp, _dp]
        L.so_residual.argtypes = [pp, _dp, _dp]
        L.so_fdjac.argtypes = [pp, _dp, cd, _dp]
        L.so_solve.argtypes = [pp, _dp, cd, ci, ctypes.POINTER(ci), _dp]
        L.so_solve.restype = ci
        L.so_traj_var.argtypes = [pp, cd, _dp, cd, _dp]
        L.so_jacobian.argtypes = [pp, _dp, _dp]
        L.so_jacobian.restype = ci
        L.so_solve_hybrj.argtypes = [pp, _dp, cd, ci, ctypes.POINTER(ci), ctypes.POINTER(ci), _dp]
        L.so_solve_hybrj.restype = ci
        L.so_continuation_param.argtypes = [pp, _dp, cd, ci, cd, ci, cd, cd, ctypes.POINTER(ci)]
        L.so_continuation_param.restype = ci
        L.so_continuation_boundary.argtypes = [pp, _dp, cd, ci, cd, _dp, ctypes.c_void_p, _dp,
                                               ctypes.c_void_p, cd, ctypes.POINTER(ci)]
        L.so_continuation_boundary.restype = ci
        _SIGS = True
    return L


def _d(a):
    return a.ctypes.data_as(_dp)


def _arr(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def load_obstacles(path):
    """Parse data/vtolUAV/obstacles the way obstacle::ReadObstacleInput does (obstacle.cpp:69-111)."""
    with open(path) as f:
        lines = [ln.strip() for ln in f.read().splitlines()]
    n = int(lines[1].split()[0])
    k = 3
    types = [float(lines[k + i].split()[0]) for i in range(n)]
    k += n + 1
    pos = [[float(v) for v in lines[k + i].split()[:3]] for i in range(n)]
    k += n + 1
    rad = [[float(v) for v in lines[k + i].split()[:3]] for i in range(n)]
    return np.array(types), np.array(pos), np.array(rad)


def make_obstacles(types, pos, rad):
    o = Obstacles()
    o.n = len(types)
    for i in range(o.n):
        o.type[i] = types[i]
        for k in range(3):
            o.pos[i][k] = pos[i][k]
            o.rad[i][k] = rad[i][k]
    return o


class OracleProblem:
    """A single OCP for the C restatement: model + shooting data (see so_problem)."""

    def __init__(self, model_id, num_multi=1, step_nbr=None, obstacles=None):
        self.p = Problem()
        lib().so_problem_init(ctypes.byref(self.p), model_id, num_multi)
        self.model_id = model_id
        self.dim = STATE_DIM[model_id]
        self.M = num_multi
        if step_nbr:
            self.p.step_nbr = step_nbr
        self._obs = obstacles
        if obstacles is not None:
            self.p.obs = ctypes.pointer(obstacles)
        # default modes (shooting::SetMode(mode_tf, mode_Xf), shooting.cpp:165-182)
        self.set_mode_final(FIXED, [FIXED] * self.dim)

    # -- model parameters -------------------------------------------------------------------
    def set_param(self, idx, value):
        self.p.mparams[idx] = value

    def get_param(self, idx):
        return self.p.mparams[idx]

    def params(self, n):
        return np.array([self.p.mparams[i] for i in range(n)])

    # -- modes / boundary data ----------------------------------------------------------------
    def set_mode_final(self, mode_tf, mode_Xf):
        M, n = self.M, self.dim
        self.p.mode_t[0] = FIXED
        for j in range(n):
            self.p.mode_X[0][j] = FIXED
        for i in range(1, M):
            self.p.mode_t[i] = CONTINUOUS
            for j in range(n):
                self.p.mode_X[i][j] = CONTINUOUS
        self.p.mode_t[M] = mode_tf
        for j in range(n):
            self.p.mode_X[M][j] = mode_Xf[j]

    def set_mode(self, mode_t, mode_X):
        for i in range(self.M + 1):
            self.p.mode_t[i] = mode_t[i]
            for j in range(self.dim):
                self.p.mode_X[i][j] = mode_X[i][j]

    def set_boundary(self, times, Xb):
        Xb = np.asarray(Xb, dtype=np.float64)
        for i in range(self.M + 1):
            self.p.time[i] = times[i]
            for j in range(self.dim):
                self.p.Xb[i][j] = Xb[i][j]

    @property
    def num_param(self):
        return lib().so_num_param(ctypes.byref(self.p))

    # -- evaluation ---------------------------------------------------------------------------
    def traj(self, t0, X0, tf):
        X0 = _arr(X0)
        out = np.zeros(2 * self.dim)
        lib().so_traj(ctypes.byref(self.p), float(t0), _d(X0), float(tf), _d(out))
        return out

    def rhs(self, t, X):
        X = _arr(X)
        out = np.zeros(2 * self.dim)
        lib().so_rhs(ctypes.byref(self.p), float(t), _d(X), _d(out))
        return out

    def control(self, t, X):
        X = _arr(X)
        out = np.zeros(4)
        n = lib().so_control(ctypes.byref(self.p), float(t), _d(X), _d(out))
        return out[:n]

    def hamiltonian(self, t, X):
        X = _arr(X)
        return lib().so_hamiltonian(ctypes.byref(self.p), float(t), _d(X))

    def obstacle(self, pos):
        pos = _arr(pos)
        f = np.zeros(1)
        g = np.zeros(3)
        lib().so_obstacle_eval(ctypes.byref(self.p), _d(pos), _d(f), _d(g))
        return f[0], g

    def timeline(self, x):
        x = _arr(x)
        tl = np.zeros(self.M + 1)
        lib().so_timeline(ctypes.byref(self.p), _d(x), _d(tl))
        return tl

    def residual(self, x):
        x = _arr(x)
        out = np.zeros(x.size)
        lib().so_residual(ctypes.byref(self.p), _d(x), _d(out))
        return out

    def fdjac(self, x, epsfcn=1e-15):
        x = _arr(x)
        out = np.zeros(x.size * x.size)
        lib().so_fdjac(ctypes.byref(self.p), _d(x), epsfcn, _d(out))
        return out.reshape(x.size, x.size).T.copy()     # J[i, j]

    def solve(self, x, xtol=1e-8, maxfev=10000):
        x = _arr(x).copy()
        nfev = ctypes.c_int(0)
        fnorm = np.zeros(1)
        info = lib().so_solve(ctypes.byref(self.p), _d(x), xtol, maxfev, ctypes.byref(nfev), _d(fnorm))
        return dict(x=x, info=info, nfev=nfev.value, fnorm=fnorm[0])

    # -- analytic-Jacobian path (modelOrder == 1: the double integrator) ------------------------
    def traj_var(self, t0, X0, tf):
        """model::ComputeTraj(isJac = 1): X0 is the (2n+1) 2n extended state."""
        X0 = _arr(X0)
        out = np.zeros(X0.size)
        lib().so_traj_var(ctypes.byref(self.p), float(t0), _d(X0), float(tf), _d(out))
        return out

    def jacobian(self, x):
        x = _arr(x)
        out = np.zeros(x.size * x.size)
        rc = lib().so_jacobian(ctypes.byref(self.p), _d(x), _d(out))
        assert rc == 0, "model has no variational equations"
        return out.reshape(x.size, x.size).T.copy()     # J[i, j]

    def solve_hybrj(self, x, xtol=1e-8, maxfev=10000):
        x = _arr(x).copy()
        nfev, njev = ctypes.c_int(0), ctypes.c_int(0)
        fnorm = np.zeros(1)
        info = lib().so_solve_hybrj(ctypes.byref(self.p), _d(x), xtol, maxfev, ctypes.byref(nfev), ctypes.byref(njev), _d(fnorm))
        return dict(x=x, info=info, nfev=nfev.value, njev=njev.value, fnorm=fnorm[0])

    def continuation_param(self, x, step, param_idx, goal, xtol=1e-8, maxfev=10000, step_min=1e-12):
        x = _arr(x).copy()
        calls = (ctypes.c_int * 2)()
        info = lib().so_continuation_param(ctypes.byref(self.p), _d(x), xtol, maxfev, step,
                                           param_idx, goal, step_min, calls)
        return dict(x=x, info=info, solver_calls=calls[0], nfev_total=calls[1])

    def continuation_boundary(self, x, step, time_prec, X_prec, timed, Xd, xtol=1e-8, maxfev=10000,
                              step_min=1e-12):
        x = _arr(x).copy()
        calls = (ctypes.c_int * 2)()

        def pad(A):
            out = np.zeros((self.M + 1, MAX_DIM))
            out[:, :self.dim] = np.asarray(A, dtype=np.float64)[:, :self.dim]
            return out
        tp, td = _arr(time_prec), _arr(timed)
        Xp, XD = pad(X_prec), pad(Xd)
        info = lib().so_continuation_boundary(ctypes.byref(self.p), _d(x), xtol, maxfev, step,
                                              _d(tp), Xp.ctypes.data, _d(td), XD.ctypes.data,
                                              step_min, calls)
        return dict(x=x, info=info, solver_calls=calls[0], nfev_total=calls[1])
