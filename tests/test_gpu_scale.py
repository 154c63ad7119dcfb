"""Batches at scale through the C ABI: the configurations of BASELINE.json beyond the bench workload
(interceptor guidance shooting, Covid-19 continuation sweep) run as real batches, every member checked
against properties that do not need the oracle at that size (identical members give identical results,
a converged member has a small residual) and a sample against the oracle."""
import numpy as np
import pytest

import scenarios as S
from golden_util import by_name, spec_from_hex
from gpu_util import engine, shape_of, batch_of

pytestmark = pytest.mark.gpu


def test_interceptor_batch_20000(oracle_lib):
    """config[2]: interceptor guidance shooting (P = 13, one warp per problem): the `initState` problem of
    tests/testInterceptor.cpp:152-165 with perturbed targets."""
    from backends import OracleBackend
    spec = spec_from_hex(by_name("solve", "interceptor_init")["spec"])
    B = 20000
    mp, time, Xb, x = batch_of([spec])
    rng = np.random.default_rng(20260003)
    Xb = np.tile(Xb, (B, 1))
    n = 6
    Xb[1:, n + 0] += rng.uniform(-500, 500, B - 1)                 # target altitude +- 500 m
    Xb[1:, n + 3] += rng.uniform(-0.05, 0.05, B - 1)               # target heading +- 0.05 rad
    Xb[B // 2] = Xb[0]                                             # a duplicate of the reference problem
    xs = np.tile(x, (B, 1))
    r = engine().solve_batch(shape_of(spec), np.tile(mp, (B, 1)), np.tile(time, (B, 1)), Xb, xs, xtol=spec["xtol"])
    o = OracleBackend().solve(spec)
    assert int(r["info"][0]) == o["info"] and int(r["info"][B // 2]) == o["info"]
    assert r["nfev"][0] == r["nfev"][B // 2] and np.array_equal(r["x"][0], r["x"][B // 2])
    ok = r["info"] == 1
    assert ok.mean() > 0.9, ok.mean()
    assert np.all(r["fnorm"][ok] < 1e-4)
    # a sample of perturbed members against the oracle
    for k in (1, 777, 19999):
        s2 = dict(spec)
        s2["Xb"] = Xb[k].reshape(2, n)
        ok_ = OracleBackend().solve(s2)
        assert int(r["info"][k]) == ok_["info"]
        if ok_["info"] == 1:
            assert np.linalg.norm(r["x"][k] - ok_["x"]) <= 10 * spec["xtol"] * np.linalg.norm(ok_["x"])


def test_covid_sweep_batch_512(oracle_lib):
    """config[4]: Covid-19 SEIR control (M = 20 segments x 1000 RK4 steps, P = 160) swept over the cost
    weights muI and Imax (the parameters the reference's demo targets, tests/testCovid19.cpp:110-115)."""
    from backends import OracleBackend
    spec = spec_from_hex(by_name("solve", "covid_stage1")["spec"])
    B = 512
    mp, time, Xb, x = batch_of([spec])
    rng = np.random.default_rng(20260005)
    mps = np.tile(mp, (B, 1))
    mps[1:, 5] = np.exp(rng.uniform(np.log(0.5), np.log(2.0), B - 1))      # muI
    mps[1:, 4] = rng.uniform(0.08, 0.12, B - 1)                            # Imax
    mps[B // 2] = mps[0]
    xs = np.tile(x, (B, 1))
    r = engine().solve_batch(shape_of(spec), mps, np.tile(time, (B, 1)), np.tile(Xb, (B, 1)), xs, xtol=spec["xtol"])
    e = by_name("solve", "covid_stage1")
    assert (int(r["info"][0]), int(r["nfev"][0])) == (e["info"], e["nfev"])                 # the reference's own result
    assert r["nfev"][0] == r["nfev"][B // 2] and np.array_equal(r["x"][0], r["x"][B // 2])
    ok = r["info"] == 1
    assert ok.mean() > 0.8, ok.mean()
    assert np.all(r["fnorm"][ok] < 1e-5)
    k = 100
    s2 = dict(spec)
    s2["mparams"] = list(mps[k])
    o = OracleBackend().solve(s2)
    assert int(r["info"][k]) == o["info"]
    if o["info"] == 1:
        assert np.linalg.norm(r["x"][k] - o["x"]) <= 10 * spec["xtol"] * np.linalg.norm(o["x"])
