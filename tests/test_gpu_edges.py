"""Edge cases through the C ABI: empty batches, zero-length and backward segments (odeTools::integrate
does nothing when tf <= t0, odeTools.cpp:135), bad arguments, NaN guesses, device-pointer calls."""
import ctypes

import numpy as np
import pytest

import scenarios as S
from golden_util import by_name, spec_from_hex
from gpu_util import engine, shape_of, batch_of

pytestmark = pytest.mark.gpu


def test_empty_batches():
    import socp_b200 as sb
    eng = engine()
    spec = S.di_problem()
    shape = shape_of(spec)
    mp, time, Xb, x = batch_of([spec])
    e = np.zeros((0, 12))
    assert eng.traj_batch(S.DI, np.zeros((0, 3)), np.zeros(0), e, np.zeros(0)).shape == (0, 12)
    assert eng.residual_batch(shape, mp[:0], time[:0], Xb[:0], x[:0]).shape == (0, 13)
    r = eng.solve_batch(shape, mp[:0], time[:0], Xb[:0], np.ascontiguousarray(x[:0]))
    assert r["info"].shape == (0,)
    r = eng.continuation_param_batch(shape, mp[:0], time[:0], Xb[:0], x[:0], 1.0, 2, np.zeros(0))
    assert r["info"].shape == (0,)


@pytest.mark.parametrize("model", range(5))
def test_zero_and_backward_segments_do_nothing(model):
    eng = engine()
    n = 2 * S.DIM[model]
    rng = np.random.default_rng(model)
    X0 = rng.uniform(0.5, 1.5, (4, n))
    if model == S.INTERCEPTOR:
        X0[:, :6] = np.array(S.INTERCEPTOR_INIT_XI)
    mp = np.array(S.DEFAULTS[model], dtype=np.float64)
    t0 = np.array([0.0, 1.0, 2.0, 50.0])
    tf = np.array([0.0, 1.0, 1.5, 30.0])                   # tf <= t0: zero steps
    if model == S.INTERCEPTOR:
        pytest.skip("interceptor::ModelInt always takes stepNbr steps (interceptor.cpp:112), also backwards")
    out = eng.traj_batch(model, mp, t0, X0, tf)
    assert np.array_equal(out, X0)
    out, ns = eng.traj_adaptive_batch(model, mp, t0, X0, tf, 1e-8)
    assert np.array_equal(out, X0) and np.all(ns == 0)


def test_bad_arguments_are_refused():
    import socp_b200 as sb
    eng = engine()
    spec = S.di_problem()
    mp, time, Xb, x = batch_of([spec])
    with pytest.raises((sb.SocpError, ValueError)):
        eng.traj_batch(7, mp, 0.0, np.zeros((1, 12)), 1.0)                       # unknown model id
    bad = shape_of(spec)
    bad.num_multi = 0
    with pytest.raises((sb.SocpError, ValueError)):
        eng.residual_batch(bad, mp, time, Xb, x)
    assert sb._lib.lib().socp_residual_batch(eng._h, ctypes.byref(bad), 1, mp.ctypes.data, time.ctypes.data, Xb.ctypes.data,
                                             x.ctypes.data, x.ctypes.data, 0) != 0          # the C ABI itself refuses it
    bad = shape_of(spec)
    bad.integrator = 1                                                            # dopri5 without a tolerance
    with pytest.raises(sb.SocpError):
        eng.residual_batch(bad, mp, time, Xb, x)
    with pytest.raises(sb.SocpError):
        eng.traj_adaptive_batch(S.DI, mp[0], 0.0, np.zeros((1, 12)), 1.0, -1.0)
    L = sb._lib.lib()
    assert L.socp_sync(None) != 0 and L.socp_num_param(None) < 0


def test_nan_guess_retires_without_success():
    """A problem whose residual is NaN must not report convergence nor stall the batch; its healthy
    neighbours are unaffected."""
    spec = spec_from_hex(by_name("solve", "di_free_tf")["spec"])
    mp, time, Xb, x = batch_of([spec] * 3)
    x = np.ascontiguousarray(x)
    x[1, 7] = np.nan
    r = engine().solve_batch(shape_of(spec), mp, time, Xb, x, xtol=spec["xtol"], maxfev=400)
    assert r["info"][0] == 1 and r["info"][2] == 1 and r["nfev"][0] == r["nfev"][2] == 82
    assert r["info"][1] != 1


def test_device_pointers_match_host_buffers():
    import torch
    import socp_b200 as sb
    spec = spec_from_hex(by_name("solve", "di_free_tf")["spec"])
    mp, time, Xb, x = batch_of([spec] * 64)
    eng = sb.Engine(0)
    host = eng.solve_batch(shape_of(spec), mp, time, Xb, np.ascontiguousarray(x).copy(), xtol=spec["xtol"])
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    dx = t(x)
    eng.use_torch_stream()
    r = eng.solve_batch(shape_of(spec), t(mp), t(time), t(Xb), dx, xtol=spec["xtol"])
    eng.sync()
    assert np.array_equal(r["x"].cpu().numpy(), host["x"])
    assert np.array_equal(r["info"].cpu().numpy(), host["info"]) and np.array_equal(r["nfev"].cpu().numpy(), host["nfev"])
