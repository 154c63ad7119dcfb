"""Helpers to push scenario specs (tests/scenarios.py) through the product's C ABI."""
import numpy as np

import scenarios as S

_ENG = None


def engine():
    global _ENG
    if _ENG is None:
        import socp_b200 as sb
        _ENG = sb.Engine(0)
        o = S.VTOL_OBSTACLES
        _ENG.set_obstacles(o["type"], o["pos"], o["rad"])
    return _ENG


def shape_of(spec):
    import socp_b200 as sb
    return sb.make_shape(spec["model"], spec["M"], spec["mode_t"], spec["mode_X"], spec["steps"])


def batch_of(specs):
    """Stack specs that share one shape into the batched arrays of the C ABI."""
    mp = np.array([s["mparams"] for s in specs], dtype=np.float64)
    time = np.array([s["time"] for s in specs], dtype=np.float64)
    Xb = np.array([np.asarray(s["Xb"], dtype=np.float64).reshape(-1) for s in specs])
    x = np.array([s["x0"] for s in specs], dtype=np.float64)
    return mp, time, Xb, x


def gpu_residual(spec, x=None):
    mp, time, Xb, x0 = batch_of([spec])
    if x is not None:
        x0 = np.asarray(x, dtype=np.float64)[None, :]
    return engine().residual_batch(shape_of(spec), mp, time, Xb, x0)[0]


def gpu_fdjac(spec, x=None):
    mp, time, Xb, x0 = batch_of([spec])
    if x is not None:
        x0 = np.asarray(x, dtype=np.float64)[None, :]
    return engine().fdjac_batch(shape_of(spec), mp, time, Xb, x0)[0]


def gpu_solve(specs, maxfev=10000):
    if isinstance(specs, dict):
        specs = [specs]
    mp, time, Xb, x = batch_of(specs)
    r = engine().solve_batch(shape_of(specs[0]), mp, time, Xb, np.ascontiguousarray(x), xtol=specs[0]["xtol"], maxfev=maxfev)
    return r
