"""GPU parity of the analytic-Jacobian path (modelOrder == 1): variational integration
(socp_traj_var_batch), shooting::ShootingFunctionJacobian (socp_jacobian_batch) and hybrj
(socp_solve_hybrj_batch) against the oracle, which is pinned bit-exact against the unmodified
reference in tests/test_oracle_hybrj.py."""
import numpy as np
import pytest

import scenarios as S
from gpu_util import engine, shape_of, batch_of
from test_oracle_hybrj import di_wp_spec

pytestmark = pytest.mark.gpu


def test_variational_trajectories(oracle_lib):
    from backends import OracleBackend
    p = OracleBackend().problem(S.di_problem())
    rng = np.random.default_rng(3)
    X0 = np.zeros((6, 156))
    X0[:, :12] = rng.uniform(-1, 1, (6, 12))
    X0[:, 12:] = np.eye(12).reshape(-1) + 0.1 * rng.uniform(-1, 1, (6, 144))
    tf = rng.uniform(1, 20, 6)
    got = engine().traj_var_batch(S.DI, np.array(S.DEFAULTS[S.DI]), 0.0, X0, tf)
    for k in range(6):
        want = p.traj_var(0.0, X0[k], tf[k])
        assert np.max(np.abs(got[k] - want)) <= 1e-12 * np.max(np.abs(want))


@pytest.mark.parametrize("make", [S.di_problem, di_wp_spec])
def test_analytic_jacobian(oracle_lib, make):
    from backends import OracleBackend
    spec = make()
    p = OracleBackend().problem(spec)
    mp, time, Xb, x = batch_of([spec])
    rng = np.random.default_rng(11)
    xs = np.vstack([x[0] * (1 + 0.05 * k * rng.uniform(-1, 1, x.shape[1])) for k in range(3)])
    got = engine().jacobian_batch(shape_of(spec), np.tile(mp, (3, 1)), np.tile(time, (3, 1)), np.tile(Xb, (3, 1)), xs)
    for k in range(3):
        want = p.jacobian(xs[k])
        # (cancelling sums of the dH rows leave rounding residue in one and an exact zero in the other)
        assert np.max(np.abs(got[k] - want)) <= 1e-12 * np.max(np.abs(want))


def test_hybrj_solves_match_reference(oracle_lib):
    """tests/testDoubleIntegrator.cpp, modelOrder = 1: nfev 30 / njev 4, and the
    two-segment waypoint problem of testDoubleIntegrator_WP.cpp."""
    from backends import OracleBackend
    for spec in (S.di_problem(), di_wp_spec()):
        p = OracleBackend().problem(spec)
        o = p.solve_hybrj(spec["x0"], xtol=spec["xtol"])
        mp, time, Xb, x = batch_of([spec] * 3)
        r = engine().solve_hybrj_batch(shape_of(spec), mp, time, Xb, np.ascontiguousarray(x), xtol=spec["xtol"])
        assert list(r["info"]) == [o["info"]] * 3
        assert list(r["nfev"]) == [o["nfev"]] * 3 and list(r["njev"]) == [o["njev"]] * 3, (r["nfev"], r["njev"], o)
        if o["info"] == 1:
            assert np.linalg.norm(r["x"][0] - o["x"]) <= spec["xtol"] * np.linalg.norm(o["x"])
    spec = S.di_problem()
    o = OracleBackend().problem(spec).solve_hybrj(spec["x0"], xtol=spec["xtol"])
    assert (o["nfev"], o["njev"]) == (30, 4)


def test_models_without_variational_equations_are_refused():
    import socp_b200 as sb
    from golden_util import by_name, spec_from_hex
    spec = spec_from_hex(by_name("solve", "covid_stage1")["spec"])
    mp, time, Xb, x = batch_of([spec])
    with pytest.raises(sb.SocpError):
        engine().solve_hybrj_batch(shape_of(spec), mp, time, Xb, np.ascontiguousarray(x))
