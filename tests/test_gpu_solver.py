"""GPU parity of the batched shooting residual, FD Jacobian and Powell-hybrid solve
(socp_residual_batch / socp_fdjac_batch / socp_solve_batch) against the golden results of the
unmodified reference and against the oracle.

Bars (BASELINE.json north_star): converged unknowns agree within the solver tolerance
(|x - x_ref| <= xtol * |x_ref|), same `info`, and the same number of residual evaluations
(nfev) on the reference's demo problems.
"""
import numpy as np
import pytest

import scenarios as S
from golden_util import golden, unhex, spec_from_hex
from gpu_util import engine, gpu_residual, gpu_fdjac, gpu_solve, shape_of, batch_of

pytestmark = pytest.mark.gpu


def test_golden_residuals():
    for e in golden()["residual"]:
        spec = spec_from_hex(e["spec"])
        want = unhex(e["fvec"])
        got = gpu_residual(spec, unhex(e["x"]))
        scale = max(np.max(np.abs(want)), np.max(np.abs(unhex(e["x"]))), 1.0)
        if spec["model"] == S.INTERCEPTOR:
            scale = max(scale, 1e3)     # altitude / velocity magnitudes enter H
        err = np.max(np.abs(got - want)) / scale
        assert err <= 1e-11, "%s residual error %.3e" % (spec["name"], err)


@pytest.mark.parametrize("name", ["di_free_tf", "goddard_stage1", "goddard_stage4_singular", "covid_stage1",
                                  "interceptor_init", "vtol_wp1"])
def test_fdjac_vs_oracle(oracle_lib, name):
    from backends import OracleBackend
    from golden_util import by_name
    spec = spec_from_hex(by_name("residual", name)["spec"])
    want = OracleBackend().fdjac(spec)
    got = gpu_fdjac(spec)
    assert got.shape == want.shape
    # structural zeros must be exact zeros
    assert np.all(got[want == 0.0] == 0.0)
    # forward differences amplify 1e-16 rounding differences by 1/h ~ 3e7/|x_j|: compare per column
    # relative to the column's largest entry
    colmax = np.maximum(np.max(np.abs(want), axis=0), 1e-300)
    err = np.max(np.abs(got - want) / colmax[None, :])
    assert err <= 1e-5, "%s FD Jacobian column-relative error %.3e" % (name, err)


DEMO_SOLVES = ["di_free_tf", "goddard_stage1", "goddard_stage4_singular", "covid_stage1", "interceptor_init",
               "vtol_wp1"]


@pytest.mark.parametrize("name", DEMO_SOLVES)
def test_demo_solves_match_reference(name):
    from golden_util import by_name
    e = by_name("solve", name)
    spec = spec_from_hex(e["spec"])
    r = gpu_solve(spec)
    xref = unhex(e["x"])
    assert r["info"][0] == e["info"] == 1
    assert np.linalg.norm(r["x"][0] - xref) <= spec["xtol"] * np.linalg.norm(xref), name
    assert r["nfev"][0] == e["nfev"], "%s: nfev %d vs reference %d" % (name, r["nfev"][0], e["nfev"])


def test_batch_members_are_independent():
    """The same problem replicated and mixed with others gives the same answer for every copy,
    and problems retire independently (convergence mask)."""
    from golden_util import by_name
    specs = [spec_from_hex(by_name("solve", "goddard_stage1")["spec"])]
    specs += [spec_from_hex(e["spec"]) for e in golden()["solve"] if e["spec"]["name"].startswith("goddard_batch")]
    specs = specs * 3
    r = gpu_solve(specs)
    k = len(specs) // 3
    for a in range(k):
        assert r["info"][a] == r["info"][a + k] == r["info"][a + 2 * k]
        assert r["nfev"][a] == r["nfev"][a + k] == r["nfev"][a + 2 * k]
        assert np.array_equal(r["x"][a], r["x"][a + k])
    solo = gpu_solve(specs[1])
    assert solo["info"][0] == r["info"][1] and solo["nfev"][0] == r["nfev"][1]
    assert np.array_equal(solo["x"][0], r["x"][1])


def test_failed_solves_report_like_minpack():
    """maxfev exhaustion gives info 2 (hybrd), and bad arguments are rejected."""
    from golden_util import by_name
    import socp_b200 as sb
    spec = spec_from_hex(by_name("solve", "goddard_stage1")["spec"])
    r = gpu_solve(spec, maxfev=200)
    assert r["info"][0] == 2 and 200 <= r["nfev"][0] < 200 + 86
    mp, time, Xb, x = batch_of([spec])
    with pytest.raises(sb.SocpError):
        engine().solve_batch(shape_of(spec), mp, time, Xb, x, xtol=-1.0)


def test_continuation_param_matches_reference():
    for e in golden()["cont_param"]:
        spec = spec_from_hex(e["spec"])
        if spec["name"] == "vtol_cont_ca":
            pass
        mp, time, Xb, x = batch_of([spec])
        r = engine().continuation_param_batch(shape_of(spec), mp, time, Xb, x, e["step"],
                                              S.pidx(spec["model"], e["pname"]), unhex(e["goal"]), xtol=spec["xtol"])
        xref = unhex(e["x"])
        assert r["info"][0] == e["info"], spec["name"]
        assert r["calls"][0, 0] == e["solver_calls"], spec["name"]
        assert np.linalg.norm(r["x"][0] - xref) <= spec["xtol"] * np.linalg.norm(xref), spec["name"]
        assert r["mparams"][0, S.pidx(spec["model"], e["pname"])] == unhex(e["goal"])
        assert abs(int(r["calls"][0, 1]) - e["nfev_total"]) <= 0.02 * e["nfev_total"] + 2, \
            "%s nfev %d vs %d" % (spec["name"], r["calls"][0, 1], e["nfev_total"])


def test_continuation_boundary_matches_reference():
    for e in golden()["cont_boundary"]:
        spec = spec_from_hex(e["spec"])
        if spec["name"] == "interceptor_S1":
            continue        # does not converge in the reference either; a chaotic homotopy
        mp, time, Xb, x = batch_of([spec])
        r = engine().continuation_boundary_batch(shape_of(spec), mp, time, Xb, unhex(e["timed"])[None, :],
                                                 unhex(e["Xd"]).reshape(1, -1), x, e["step"], xtol=spec["xtol"])
        xref = unhex(e["x"])
        assert r["info"][0] == e["info"] == 1, spec["name"]
        assert r["calls"][0, 0] == e["solver_calls"], spec["name"]
        assert np.linalg.norm(r["x"][0] - xref) <= spec["xtol"] * np.linalg.norm(xref), spec["name"]
