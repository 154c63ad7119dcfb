"""Pin the clean-room MINPACK restatement (oracle/minpack.c) against scipy's MINPACK.

The reference solves with cminpack's hybrd/hybrj (/root/reference/src/socp/shooting.cpp:803-851);
cminpack is not vendored.  scipy.optimize._minpack wraps the original MINPACK Fortran (same
algorithm lineage) and exposes every knob SOCP sets (xtol, maxfev, ml, mu, epsfcn, factor, diag),
so the pin is: same x (bit for bit), same info, same nfev/njev on standard systems.
"""
import numpy as np
import pytest
from scipy.optimize import _minpack as sm

from oracle import pyminpack as pm


def rosen(x):
    return np.array([10 * (x[1] - x[0] ** 2), 1 - x[0]])


def rosen_jac(x):
    return np.array([[-20 * x[0], 10.0], [-1.0, 0.0]])


def powell_sing(x):
    return np.array([x[0] + 10 * x[1], np.sqrt(5) * (x[2] - x[3]), (x[1] - 2 * x[2]) ** 2,
                     np.sqrt(10) * (x[0] - x[3]) ** 2])


def powell_sing_jac(x):
    return np.array([[1, 10, 0, 0], [0, 0, np.sqrt(5), -np.sqrt(5)],
                     [0, 2 * (x[1] - 2 * x[2]), -4 * (x[1] - 2 * x[2]), 0],
                     [2 * np.sqrt(10) * (x[0] - x[3]), 0, 0, -2 * np.sqrt(10) * (x[0] - x[3])]])


def broyden_tri(x):
    n = len(x)
    f = np.zeros(n)
    for k in range(n):
        t1 = x[k - 1] if k > 0 else 0.0
        t2 = x[k + 1] if k < n - 1 else 0.0
        f[k] = (3 - 2 * x[k]) * x[k] - t1 - 2 * t2 + 1
    return f


def broyden_tri_jac(x):
    n = len(x)
    J = np.zeros((n, n))
    for k in range(n):
        J[k, k] = 3 - 4 * x[k]
        if k > 0:
            J[k, k - 1] = -1
        if k < n - 1:
            J[k, k + 1] = -2
    return J


def helical(x):
    tpi = 8 * np.arctan(1.0)
    t = 0.25 if x[1] >= 0 else -0.25
    if x[0] > 0:
        t = np.arctan(x[1] / x[0]) / tpi
    if x[0] < 0:
        t = np.arctan(x[1] / x[0]) / tpi + 0.5
    return np.array([10 * (x[2] - 10 * t), 10 * (np.sqrt(x[0] ** 2 + x[1] ** 2) - 1), x[2]])


def trig(x):
    n = len(x)
    return n - np.sum(np.cos(x)) + np.arange(1, n + 1) * (1 - np.cos(x)) - np.sin(x)


CASES = [(rosen, [-1.2, 1.0]), (rosen, [-12.0, 10.0]), (powell_sing, [3, -1, 0, 1.0]),
         (powell_sing, [30, -10, 0, 10.0]), (broyden_tri, -np.ones(10)),
         (broyden_tri, -np.ones(40)), (helical, [-1, 0, 0.0]), (trig, np.ones(10) / 10)]
# (xtol, epsfcn, factor): first row = SOCP's defaults (shooting.cpp:96-100), second = testGoddard's
SETTINGS = [(1e-8, 1e-15, 1.0), (1e-6, 1e-15, 1.0), (1e-10, 0.0, 100.0)]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("setting", SETTINGS)
def test_hybrd_matches_scipy_minpack(oracle_lib, case, setting):
    f, x0 = CASES[case]
    xtol, eps, fac = setting
    n = len(x0)
    mine = pm.hybrd(f, x0, xtol=xtol, epsfcn=eps, factor=fac)
    x, d, info = sm._hybrd(f, np.array(x0, float), (), 1, xtol, 10000, n - 1, n - 1, eps, fac, None)
    assert mine["info"] == info
    assert mine["nfev"] == d["nfev"]
    assert np.array_equal(mine["x"], x)          # bit-exact iterates
    assert np.array_equal(mine["fvec"], d["fvec"])


def test_hybrd_banded_matches_scipy(oracle_lib):
    x0 = -np.ones(30)
    mine = pm.hybrd(broyden_tri, x0, ml=1, mu=1)
    x, d, info = sm._hybrd(broyden_tri, x0.copy(), (), 1, 1e-8, 10000, 1, 1, 1e-15, 1.0, None)
    assert (mine["info"], mine["nfev"]) == (info, d["nfev"])
    assert np.array_equal(mine["x"], x)


JCASES = [(rosen, rosen_jac, [-1.2, 1.0]), (powell_sing, powell_sing_jac, [3, -1, 0, 1.0]),
          (broyden_tri, broyden_tri_jac, -np.ones(12))]


@pytest.mark.parametrize("case", range(len(JCASES)))
@pytest.mark.parametrize("factor", [1.0, 100.0])
def test_hybrj_matches_scipy_minpack(oracle_lib, case, factor):
    f, jac, x0 = JCASES[case]
    mine = pm.hybrj(f, jac, x0, factor=factor)
    x, d, info = sm._hybrj(f, jac, np.array(x0, float), (), 1, 0, 1e-8, 10000, factor, None)
    assert mine["info"] == info
    assert (mine["nfev"], mine["njev"]) == (d["nfev"], d["njev"])
    assert np.array_equal(mine["x"], x)


def test_enorm_extremes(oracle_lib):
    import ctypes
    for v in ([3e-200, 4e-200], [3e200, 4e200], [1e-30, 1.0, 1e30], [0.0, 0.0]):
        a = np.array(v)
        got = oracle_lib.mp_enorm(len(v), a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        want = float(np.linalg.norm(a / max(np.max(np.abs(a)), 1e-300)) * max(np.max(np.abs(a)), 1e-300)) if np.any(a) else 0.0
        assert got == pytest.approx(want, rel=1e-14)


def test_bad_input_returns_info0(oracle_lib):
    assert pm.hybrd(rosen, [1.0, 1.0], xtol=-1.0)["info"] == 0
    assert pm.hybrd(rosen, [1.0, 1.0], factor=0.0)["info"] == 0
