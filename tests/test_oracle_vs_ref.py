"""Direct pin of the C restatement against the unmodified reference build (oracle/_ref), on
seeded random inputs beyond the committed fixtures.  Skipped where the prebuilt _ref library is
absent (it is built by `make -C oracle ref` only where /root/reference exists)."""
import numpy as np
import pytest

import scenarios as S
from oracle import pyref as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libsocp_ref.so not built")


def _backends():
    from backends import OracleBackend, RefBackend
    return OracleBackend(), RefBackend()


@pytest.mark.parametrize("model", [S.GODDARD, S.DI, S.COVID19, S.VTOL, S.INTERCEPTOR])
def test_random_trajectories_bit_exact(oracle_lib, model):
    ora, ref = _backends()
    rng = np.random.default_rng(20260100 + model)
    base = {S.GODDARD: S.GODDARD_XI, S.DI: np.r_[np.zeros(6), 0.01 * np.ones(6)], S.COVID19: S.COVID_XI,
            S.VTOL: np.array([20, 8, 5, 0.3, 0.2, 0.1, -0.03, 0.013, -0.003, -0.29, 0.1, -0.03]),
            S.INTERCEPTOR: np.array(S.INTERCEPTOR_INIT_XI + [0.01, -1, 0.5, 0.2, 100., 50.])}[model]
    mp = list(S.DEFAULTS[model])
    if model == S.GODDARD:
        mp[6] = 1.0
    tf = {S.GODDARD: 0.05, S.DI: 8.0, S.COVID19: 20.0, S.VTOL: 5.0, S.INTERCEPTOR: 25.0}[model]
    for _ in range(8):
        X0 = base * (1.0 + 0.05 * rng.uniform(-1, 1, size=base.size)) + 1e-3 * rng.uniform(-1, 1, size=base.size)
        a = ora.traj(model, mp, 0.0, X0, tf)
        b = ref.traj(model, mp, 0.0, X0, tf)
        assert np.array_equal(a, b)


def test_goddard_residual_and_solve_bit_exact(oracle_lib):
    ora, ref = _backends()
    Xi, xf = S.goddard_batch_inputs(2, seed=77)
    for k in range(2):
        spec = S.goddard_problem(lambda mp, a, b, c: ora.traj(S.GODDARD, mp, a, b, c, 10), Xi=Xi[k], xf0=xf[k])
        assert np.array_equal(ora.residual(spec), ref.residual(spec))
        ro, rr = ora.solve(spec), ref.solve(spec)
        assert (ro["info"], ro["nfev"]) == (rr["info"], rr["nfev"])
        if rr["info"] == 1:
            assert np.array_equal(ro["x"], rr["x"])
