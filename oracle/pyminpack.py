"""ctypes front-end to the clean-room MINPACK restatement (oracle/minpack.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's cpu_baseline / reference arm and
__graft_entry__.smoke(); never by the socp_b200 package.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FUNC_NN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                           ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                           ctypes.c_int)
FUNCDER_NN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                              ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                              ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int)


def build(force=False):
    """Compile oracle/*.c into oracle/liboracle.so (gcc, no FMA contraction)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"] + (["-B"] if force else []))
    return os.path.join(_HERE, "liboracle.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        dp = ctypes.POINTER(ctypes.c_double)
        _LIB.hybrd.restype = ctypes.c_int
        _LIB.hybrd.argtypes = [FUNC_NN, ctypes.c_void_p, ctypes.c_int, dp, dp, ctypes.c_double,
                               ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, dp,
                               ctypes.c_int, ctypes.c_double, ctypes.c_int,
                               ctypes.POINTER(ctypes.c_int), dp, ctypes.c_int, dp, ctypes.c_int,
                               dp, dp, dp, dp, dp]
        _LIB.hybrj.restype = ctypes.c_int
        _LIB.hybrj.argtypes = [FUNCDER_NN, ctypes.c_void_p, ctypes.c_int, dp, dp, dp, ctypes.c_int,
                               ctypes.c_double, ctypes.c_int, dp, ctypes.c_int, ctypes.c_double,
                               ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                               ctypes.POINTER(ctypes.c_int), dp, ctypes.c_int, dp, dp, dp, dp, dp]
        _LIB.mp_enorm.restype = ctypes.c_double
        _LIB.mp_enorm.argtypes = [ctypes.c_int, dp]
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def hybrd(func, x0, xtol=1e-8, maxfev=10000, ml=None, mu=None, epsfcn=1e-15, factor=1.0,
          mode=1, diag=None):
    """Solve func(x)=0 with the oracle hybrd.  Defaults are SOCP's (shooting.cpp:95-101)."""
    x = np.array(x0, dtype=np.float64).copy()
    n = x.size
    ml = n - 1 if ml is None else ml
    mu = n - 1 if mu is None else mu

    def cb(_p, nn, xp, fp, _iflag):
        xv = np.ctypeslib.as_array(xp, shape=(nn,))
        fv = np.ctypeslib.as_array(fp, shape=(nn,))
        fv[:] = func(xv.copy())
        return 0

    fvec = np.zeros(n)
    d = np.ones(n) if diag is None else np.array(diag, dtype=np.float64)
    fjac = np.zeros(n * n)
    lr = n * (n + 1) // 2
    r = np.zeros(lr)
    qtf = np.zeros(n)
    wa = [np.zeros(n) for _ in range(4)]
    nfev = ctypes.c_int(0)
    info = lib().hybrd(FUNC_NN(cb), None, n, _dp(x), _dp(fvec), xtol, maxfev, ml, mu, epsfcn,
                       _dp(d), mode, factor, 0, ctypes.byref(nfev), _dp(fjac), n, _dp(r), lr,
                       _dp(qtf), _dp(wa[0]), _dp(wa[1]), _dp(wa[2]), _dp(wa[3]))
    return dict(x=x, fvec=fvec, info=info, nfev=nfev.value, fjac=fjac.reshape(n, n), r=r, qtf=qtf)


def hybrj(func, jac, x0, xtol=1e-8, maxfev=10000, factor=1.0, mode=1, diag=None):
    """Solve func(x)=0 with analytic Jacobian jac(x) -> (n,n) array J[i,j]=dF_i/dx_j."""
    x = np.array(x0, dtype=np.float64).copy()
    n = x.size

    def cb(_p, nn, xp, fp, jp, ld, iflag):
        xv = np.ctypeslib.as_array(xp, shape=(nn,)).copy()
        if iflag == 1:
            np.ctypeslib.as_array(fp, shape=(nn,))[:] = func(xv)
        else:
            J = np.asarray(jac(xv), dtype=np.float64)
            np.ctypeslib.as_array(jp, shape=(nn * ld,))[:] = J.T.reshape(-1)  # column-major
        return 0

    fvec = np.zeros(n)
    d = np.ones(n) if diag is None else np.array(diag, dtype=np.float64)
    fjac = np.zeros(n * n)
    lr = n * (n + 1) // 2
    r = np.zeros(lr)
    qtf = np.zeros(n)
    wa = [np.zeros(n) for _ in range(4)]
    nfev = ctypes.c_int(0)
    njev = ctypes.c_int(0)
    info = lib().hybrj(FUNCDER_NN(cb), None, n, _dp(x), _dp(fvec), _dp(fjac), n, xtol, maxfev,
                       _dp(d), mode, factor, 0, ctypes.byref(nfev), ctypes.byref(njev), _dp(r),
                       lr, _dp(qtf), _dp(wa[0]), _dp(wa[1]), _dp(wa[2]), _dp(wa[3]))
    return dict(x=x, fvec=fvec, info=info, nfev=nfev.value, njev=njev.value)
