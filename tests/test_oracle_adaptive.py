"""CPU anchors of the oracle's adaptive Dormand-Prince restatement (oracle/socp_oracle.c,
so_integrate_adaptive; the reference's -D_USE_BOOST build, odeTools.cpp:131-134).

Boost.Odeint is neither in the reference tree nor in this image, so the controller stays PARITY UNPINNED
against Boost itself.  What can be pinned independently is pinned here against scipy's RK45, which is an
independent implementation of the same Dormand-Prince 5(4) pair:
  * the tableau and stage times: ONE step of the oracle equals scipy's rk_step to rounding;
  * the embedded error estimate: the tolerance at which the oracle starts accepting that step is the one
    scipy's error coefficients give for Boost's error norm  |err_i| / (tol + tol (|x_i| + dt |dxdt_i|));
  * the end point of a whole integration agrees with scipy's solve_ivp within the requested tolerance
    (the two step-size controllers differ, the integrated ODE and the method do not)."""
import numpy as np
import pytest

import scenarios as S

CASES = {
    S.GODDARD: (S.GODDARD_XI, 0.0, 0.1, {6: 1.0, 2: 0.0}),
    S.DI: (np.r_[np.zeros(6), 0.01 * np.ones(6)], 0.0, 8.0, {}),
    S.COVID19: (S.COVID_XI, 0.0, 30.0, {}),
}


def problem(model, over, steps, tol):
    import oracle.pyoracle as O
    p = O.OracleProblem(model, 1, step_nbr=steps)
    mp = np.array(S.DEFAULTS[model], dtype=np.float64)
    for k, v in over.items():
        mp[k] = v
    for k, v in enumerate(mp):
        p.set_param(k, v)
    p.p.integrator, p.p.ode_tol = 1, tol
    return p


def scipy_step(p, t0, X0, h):
    from scipy.integrate._ivp.rk import RK45, rk_step
    fun = lambda t, y: p.rhs(t, y)
    f0 = fun(t0, X0)
    K = np.empty((RK45.n_stages + 1, X0.size))
    y1, f1 = rk_step(fun, t0, X0, f0, h, RK45.A, RK45.B, RK45.C, K)
    err = h * (K.T @ RK45.E)
    return y1, f0, err


@pytest.mark.parametrize("model", sorted(CASES))
def test_one_step_equals_scipy_rk_step(oracle_lib, model):
    X0, t0, tf, over = CASES[model]
    h = (tf - t0) / 50
    p = problem(model, over, 1, 1e30)          # one step of length h, always accepted
    got = p.traj(t0, X0, t0 + h)
    assert (p.p.dopri_steps, p.p.dopri_rejected) == (1, 0)
    want, _, _ = scipy_step(p, t0, np.asarray(X0, dtype=np.float64), h)
    assert np.max(np.abs(got - want)) <= 1e-14 * max(1.0, np.max(np.abs(want)))


# (the double integrator's flow is piecewise quadratic: its error estimate is rounding noise, nothing to pin)
@pytest.mark.parametrize("model", [S.GODDARD, S.COVID19])
def test_error_estimate_equals_scipy_coefficients(oracle_lib, model):
    X0, t0, tf, over = CASES[model]
    X0 = np.asarray(X0, dtype=np.float64)
    h = (tf - t0) / 4
    p = problem(model, over, 1, 1.0)
    _, f0, err = scipy_step(p, t0, X0, h)
    # Boost's default_error_checker: the step is accepted iff max_i |err_i| / (tol (1 + |x_i| + h |f_i|)) <= 1
    tol_star = np.max(np.abs(err) / (1.0 + np.abs(X0) + h * np.abs(f0)))
    assert tol_star > 0

    def first_try_rejected(tol):
        q = problem(model, over, 1, tol)
        q.traj(t0, X0, t0 + h)
        return q.p.dopri_rejected > 0
    assert not first_try_rejected(tol_star * (1 + 1e-6))
    assert first_try_rejected(tol_star * (1 - 1e-6))


@pytest.mark.parametrize("model", sorted(CASES))
@pytest.mark.parametrize("tol", [1e-6, 1e-9])
def test_end_point_agrees_with_solve_ivp(oracle_lib, model, tol):
    from scipy.integrate import solve_ivp
    X0, t0, tf, over = CASES[model]
    X0 = np.asarray(X0, dtype=np.float64)
    p = problem(model, over, S.STEPS[model], tol)
    got = p.traj(t0, X0, tf)
    assert p.p.dopri_steps >= 1
    ref = solve_ivp(lambda t, y: p.rhs(t, y), (t0, tf), X0, method="DOP853", rtol=1e-13, atol=1e-13).y[:, -1]
    sci = solve_ivp(lambda t, y: p.rhs(t, y), (t0, tf), X0, method="RK45", rtol=tol, atol=tol).y[:, -1]
    scale = max(1.0, np.max(np.abs(ref)))
    e_oracle = np.max(np.abs(got - ref)) / scale
    e_scipy = np.max(np.abs(sci - ref)) / scale
    # global error of either controller is a modest multiple of the local tolerance; the two must be
    # of the same order (Boost's controller is not less accurate than scipy's by more than 30x)
    assert e_oracle <= 300 * tol, (e_oracle, e_scipy)
    assert e_oracle <= 30 * max(e_scipy, tol), (e_oracle, e_scipy)
