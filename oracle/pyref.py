"""ctypes front-end to oracle/_ref/libsocp_ref.so: the UNMODIFIED reference (bherisse/socp)
compiled from /root/reference by oracle/Makefile, with oracle/minpack.c behind hybrd/hybrj.

TEST INFRASTRUCTURE ONLY (tests/, bench.py cpu_baseline / --impl reference, smoke()).
The .so is built in the authoring container and travels to the GPU box; /root/reference is
never read at run time.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_ref", "libsocp_ref.so")

GODDARD, DI, COVID19, VTOL, INTERCEPTOR = range(5)
MODEL_NAMES = ["goddard", "doubleIntegrator", "covid19", "vtolUAV", "interceptor"]
STATE_DIM = {GODDARD: 7, DI: 6, COVID19: 4, VTOL: 6, INTERCEPTOR: 6}
# parameter index maps (shared with the product's parameter blocks, see include/socp_b200.h)
PARAMS = {
    GODDARD: ["C", "b", "KD", "kr", "u_max", "mu1", "mu2", "singularControl"],
    DI: ["u_max", "a_max", "muT"],
    COVID19: ["R0", "Tinf", "Tinc", "N", "Imax", "muI", "umin", "umax"],
    VTOL: ["u_max", "a_max", "alphaT", "alphaV", "invSigmaXwp", "Vd", "ca", "nWP_tot", "nWP",
           "phiObs", "psiWP", "muObs", "sigmaWP"],
    INTERCEPTOR: ["c0", "hr", "d0", "eta", "propellant_mass", "empty_mass", "q", "ve", "alpha_max",
                  "u_max", "a_max", "r_2p", "t_2p", "mu_gft", "muT", "muV", "muC"],
}
FIXED, FREE, CONTINUOUS = 0, 1, 2

_LIB = None
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def available():
    return os.path.exists(SO_PATH)


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(SO_PATH)
        vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        sig = {
            "ref_model_new": (vp, [ci, ci, ci, ctypes.c_char_p, ctypes.c_char_p]),
            "ref_model_free": (None, [vp]),
            "ref_model_dim": (ci, [vp]),
            "ref_model_set_param": (ci, [vp, ci, cd]),
            "ref_model_get_param": (cd, [vp, ci]),
            "ref_model_set_ode_tol": (None, [vp, cd]),
            "ref_model_switching_times": (None, [vp, ci, _dp]),
            "ref_traj": (None, [vp, cd, _dp, ci, cd, ci, _dp]),
            "ref_rhs": (ci, [vp, cd, _dp, ci, ci, _dp]),
            "ref_control": (ci, [vp, cd, _dp, ci, _dp]),
            "ref_hamiltonian": (ci, [vp, cd, _dp, ci, ci, _dp]),
            "ref_obstacle_eval": (None, [vp, _dp, _dp, _dp]),
            "ref_interceptor_init_analytical": (None, [vp, cd, _dp, cd, _dp]),
            "ref_shooting_new": (vp, [vp, ci, ci]),
            "ref_shooting_free": (None, [vp]),
            "ref_shooting_resize": (None, [vp, ci, ci]),
            "ref_shooting_set_precision": (None, [vp, cd]),
            "ref_shooting_set_cont_min_step": (None, [vp, cd]),
            "ref_shooting_set_mode_final": (None, [vp, ci, _ip, ci]),
            "ref_shooting_set_mode": (None, [vp, _ip, _ip, ci, ci]),
            "ref_shooting_init": (None, [vp, cd, _dp, cd, _dp, ci]),
            "ref_shooting_init_v": (None, [vp, _dp, _dp, ci, ci]),
            "ref_shooting_desired": (None, [vp, cd, _dp, cd, _dp, ci]),
            "ref_shooting_desired_v": (None, [vp, _dp, _dp, ci, ci]),
            "ref_shooting_solve": (ci, [vp, cd]),
            "ref_shooting_solve_param": (ci, [vp, vp, cd, ci, cd]),
            "ref_shooting_num_param": (ci, [vp]),
            "ref_shooting_get_params": (None, [vp, _dp]),
            "ref_shooting_get_solution": (None, [vp, ci, ci, _dp, _dp]),
            "ref_shooting_move": (None, [vp, cd, ci, _dp]),
            "ref_shooting_call_number": (None, [vp, _ip]),
            "ref_shooting_residual": (None, [vp, _dp, _dp]),
            "ref_shooting_jacobian": (None, [vp, _dp, _dp]),
            "ref_log_clear": (None, []),
            "ref_log_size": (ci, []),
            "ref_log_get": (None, [ci, _ip]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _arr(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


class RefModel:
    """One of the five reference models (src/models/*), constructed with an empty trace path."""

    def __init__(self, model_id, model_order=0, step_nbr=0, obstacle_file=None, wp_file=None):
        self.id = model_id
        self.h = lib().ref_model_new(model_id, model_order, step_nbr,
                                     obstacle_file.encode() if obstacle_file else None,
                                     wp_file.encode() if wp_file else None)
        self.dim = lib().ref_model_dim(self.h)
        self.model_order = model_order

    def set(self, name, value):
        rc = lib().ref_model_set_param(self.h, PARAMS[self.id].index(name), float(value))
        assert rc == 0

    def get(self, name):
        return lib().ref_model_get_param(self.h, PARAMS[self.id].index(name))

    def params(self):
        return np.array([lib().ref_model_get_param(self.h, k) for k in range(len(PARAMS[self.id]))])

    def switching_times(self, ts):
        ts = _arr(ts)
        lib().ref_model_switching_times(self.h, ts.size, _d(ts))

    def traj(self, t0, X0, tf, is_jac=0):
        """model::ComputeTraj (model.hpp:77)."""
        X0 = _arr(X0)
        out = np.zeros(X0.size)
        lib().ref_traj(self.h, float(t0), _d(X0), X0.size, float(tf), is_jac, _d(out))
        return out

    def rhs(self, t, X, is_jac=0):
        X = _arr(X)
        out = np.zeros(max(X.size, 4 * self.dim * self.dim + 2 * self.dim))
        n = lib().ref_rhs(self.h, float(t), _d(X), X.size, is_jac, _d(out))
        return out[:n]

    def control(self, t, X):
        X = _arr(X)
        out = np.zeros(8)
        n = lib().ref_control(self.h, float(t), _d(X), X.size, _d(out))
        return out[:n]

    def hamiltonian(self, t, X, is_jac=0):
        X = _arr(X)
        out = np.zeros(2 * self.dim + 1)
        n = lib().ref_hamiltonian(self.h, float(t), _d(X), X.size, is_jac, _d(out))
        return out[:n]

    def obstacle(self, pos):
        pos = _arr(pos)
        f = np.zeros(1)
        g = np.zeros(3)
        lib().ref_obstacle_eval(self.h, _d(pos), _d(f), _d(g))
        return f[0], g

    def init_analytical(self, ti, Xi, tf, Xf):
        Xi, Xf = _arr(Xi).copy(), _arr(Xf).copy()
        lib().ref_interceptor_init_analytical(self.h, float(ti), _d(Xi), float(tf), _d(Xf))
        return Xi, Xf


class RefShooting:
    """The reference `shooting` object (src/socp/shooting.hpp:18)."""

    def __init__(self, model, num_multi=1, num_thread=1):
        self.model = model
        self.M = num_multi
        self.h = lib().ref_shooting_new(model.h, num_multi, num_thread)

    def resize(self, num_multi, num_thread=1):
        self.M = num_multi
        lib().ref_shooting_resize(self.h, num_multi, num_thread)

    def set_precision(self, xtol):
        lib().ref_shooting_set_precision(self.h, float(xtol))

    def set_mode_final(self, mode_tf, mode_Xf):
        m = np.ascontiguousarray(mode_Xf, dtype=np.int32)
        lib().ref_shooting_set_mode_final(self.h, int(mode_tf), _i(m), m.size)

    def set_mode(self, mode_t, mode_X):
        mt = np.ascontiguousarray(mode_t, dtype=np.int32)
        mx = np.ascontiguousarray(mode_X, dtype=np.int32)
        lib().ref_shooting_set_mode(self.h, _i(mt), _i(mx), mt.size, mx.shape[1])

    def init(self, ti, Xi, tf, Xf):
        Xi, Xf = _arr(Xi), _arr(Xf)
        lib().ref_shooting_init(self.h, float(ti), _d(Xi), float(tf), _d(Xf), Xi.size)

    def init_v(self, vt, vX):
        vt, vX = _arr(vt), _arr(vX)
        lib().ref_shooting_init_v(self.h, _d(vt), _d(vX), vt.size, vX.shape[1])

    def desired(self, ti, Xi, tf, Xf):
        Xi, Xf = _arr(Xi), _arr(Xf)
        lib().ref_shooting_desired(self.h, float(ti), _d(Xi), float(tf), _d(Xf), Xi.size)

    def desired_v(self, vt, vX):
        vt, vX = _arr(vt), _arr(vX)
        lib().ref_shooting_desired_v(self.h, _d(vt), _d(vX), vt.size, vX.shape[1])

    def solve(self, step=0.0):
        return lib().ref_shooting_solve(self.h, float(step))

    def solve_param(self, step, name, goal):
        return lib().ref_shooting_solve_param(self.h, self.model.h, float(step),
                                              PARAMS[self.model.id].index(name), float(goal))

    @property
    def num_param(self):
        return lib().ref_shooting_num_param(self.h)

    def params(self):
        out = np.zeros(self.num_param)
        lib().ref_shooting_get_params(self.h, _d(out))
        return out

    def solution(self):
        nX = 2 * self.model.dim
        vt = np.zeros(self.M + 1)
        vX = np.zeros((self.M + 1, nX))
        lib().ref_shooting_get_solution(self.h, self.M + 1, nX, _d(vt), _d(vX))
        return vt, vX

    def move(self, tf):
        out = np.zeros(2 * self.model.dim)
        lib().ref_shooting_move(self.h, float(tf), out.size, _d(out))
        return out

    def call_number(self):
        out = np.zeros(2, dtype=np.int32)
        lib().ref_shooting_call_number(self.h, _i(out))
        return int(out[0]), int(out[1])

    def residual(self, x):
        """F(x) through shooting::StaticShootingFunction with boundary data = desired data."""
        x = _arr(x)
        out = np.zeros(x.size)
        lib().ref_shooting_residual(self.h, _d(x), _d(out))
        return out

    def jacobian(self, x):
        """Analytic dF/dx (modelOrder==1), returned as J[i, j] = dF_i/dx_j."""
        x = _arr(x)
        out = np.zeros(x.size * x.size)
        lib().ref_shooting_jacobian(self.h, _d(x), _d(out))
        return out.reshape(x.size, x.size).T.copy()


def log_clear():
    lib().ref_log_clear()


def log():
    """[(info, nfev, njev, n)] for every hybrd/hybrj call since log_clear()."""
    out = []
    buf = np.zeros(4, dtype=np.int32)
    for k in range(lib().ref_log_size()):
        lib().ref_log_get(k, _i(buf))
        out.append(tuple(int(v) for v in buf))
    return out
