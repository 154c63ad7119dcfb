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


def test_two_engines_in_one_process_each_configure_their_own_kernels():
    """Opt-in shared memory (> 48 KB) is a per-device function attribute and is tracked per context: a second
    engine -- on the second GPU when there is one, else on the same GPU -- runs the P = 85 Powell-hybrid kernels
    (67-96 KB of dynamic shared memory) and gets the first engine's results bit for bit."""
    import torch
    import socp_b200 as sb
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    e0 = sb.Engine(0)
    e1 = sb.Engine(1 if torch.cuda.device_count() > 1 else 0)
    w = bench.wl_goddard_warm(e0, 64, seed=20260002)
    out = []
    for e in (e0, e1, e0):
        x = np.ascontiguousarray(w.x0).copy()
        r = e.solve_batch(w.shape, w.mp, w.time, w.Xb, x, xtol=w.xtol)
        out.append((r["x"].copy(), r["info"].copy(), r["nfev"].copy()))
    for a in out[1:]:
        assert np.array_equal(a[0], out[0][0]) and np.array_equal(a[1], out[0][1]) and np.array_equal(a[2], out[0][2])
    assert np.all(out[0][1] == 1)


def test_device_tensors_are_validated():
    """DEVICE-mode arguments are checked before their pointer crosses the C ABI: wrong dtype, wrong GPU."""
    import torch
    import socp_b200 as sb
    e = sb.Engine(0)
    X0 = torch.zeros((4, 12), dtype=torch.float32, device="cuda:0")
    mp = torch.ones((4, 3), dtype=torch.float64, device="cuda:0")
    t = torch.zeros(4, dtype=torch.float64, device="cuda:0")
    with pytest.raises(ValueError):
        e.traj_batch(sb.DOUBLE_INTEGRATOR, mp, t, X0, t + 1.0)


def test_device_mode_follows_torch_current_stream():
    """A DEVICE-mode call is ordered after work queued on torch's current stream without the caller doing
    anything: the copy that fills x on a side stream must be seen by the solve."""
    import torch
    import socp_b200 as sb
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    e = sb.Engine(0)
    w = bench.wl_goddard_warm(e, 2048, seed=20260002)
    dev = torch.device("cuda", 0)
    d = bench.device_arrays(torch, dev, w)
    ref = e.solve_batch(w.shape, w.mp, w.time, w.Xb, np.ascontiguousarray(w.x0).copy(), xtol=w.xtol)
    side = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        d["x"].zero_()
        for _ in range(20):                       # keep the side stream busy before the copy that matters
            d["x"].add_(1.0)
        d["x"].copy_(d["x0"])
        e.solve_batch(w.shape, d["mp"], d["time"], d["Xb"], d["x"], xtol=w.xtol, info=d["info"], nfev=d["nfev"], fnorm=d["fnorm"])
    torch.cuda.synchronize()
    assert np.array_equal(d["info"].cpu().numpy(), ref["info"]) and np.array_equal(d["nfev"].cpu().numpy(), ref["nfev"])
    assert np.array_equal(d["x"].cpu().numpy(), ref["x"])
