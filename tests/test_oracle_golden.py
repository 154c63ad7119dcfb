"""The plain-C restatement (oracle/socp_oracle.c) against the golden vectors that
tests/golden/make_golden.py recorded from the UNMODIFIED reference build (oracle/_ref).

Integer results (info, nfev, solver-call counts) must be identical; floating-point results are
compared bit for bit (same compiler flags, same libm in this image), falling back to a 1e-13
relative bound only if the host libm differs.
"""
import numpy as np
import pytest

import scenarios as S
from backends import OracleBackend
from golden_util import golden, unhex, spec_from_hex

ORA = OracleBackend()


def assert_same(got, want, what):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    if np.array_equal(got, want):
        return
    scale = max(np.max(np.abs(want)), 1e-300)
    err = np.max(np.abs(got - want)) / scale
    assert err <= 1e-13, "%s: not bit-exact and norm-relative error %.3e" % (what, err)


@pytest.mark.parametrize("k", range(len(golden()["traj"])))
def test_trajectory(oracle_lib, k):
    e = golden()["traj"][k]
    got = ORA.traj(e["model"], unhex(e["mparams"]), unhex(e["t0"]), unhex(e["X0"]), unhex(e["tf"]),
                   e["steps"], unhex(e["sw"]))
    assert_same(got, unhex(e["Xf"]), "traj %d model %d" % (k, e["model"]))


@pytest.mark.parametrize("k", range(len(golden()["points"])))
def test_rhs_control_hamiltonian(oracle_lib, k):
    import ctypes
    from oracle import pyoracle as O
    e = golden()["points"][k]
    p = O.OracleProblem(e["model"], 1, obstacles=None)
    if e["model"] == S.VTOL:
        from backends import oracle_obstacles
        p = O.OracleProblem(e["model"], 1, obstacles=oracle_obstacles())
    for i, v in enumerate(unhex(e["mparams"])):
        p.set_param(i, v)
    if e["sw"] is not None:
        for i, v in enumerate(unhex(e["sw"])):
            p.p.sw[i] = v
    t, X = unhex(e["t"]), unhex(e["X"])
    assert_same(p.rhs(t, X), unhex(e["rhs"]), "rhs")
    assert_same(p.control(t, X), unhex(e["control"]), "control")
    assert_same(p.hamiltonian(t, X), unhex(e["H"]), "H")


@pytest.mark.parametrize("k", range(len(golden()["residual"])))
def test_residual(oracle_lib, k):
    e = golden()["residual"][k]
    spec = spec_from_hex(e["spec"])
    assert_same(ORA.residual(spec, unhex(e["x"])), unhex(e["fvec"]), "residual " + spec["name"])


@pytest.mark.parametrize("k", range(len(golden()["solve"])))
def test_solve(oracle_lib, k):
    e = golden()["solve"][k]
    spec = spec_from_hex(e["spec"])
    r = ORA.solve(spec)
    assert (r["info"], r["nfev"]) == (e["info"], e["nfev"]), spec["name"]
    if e["info"] == 1:      # on failure the reference keeps its guess (shooting.cpp:588)
        assert_same(r["x"], unhex(e["x"]), "solve " + spec["name"])


@pytest.mark.parametrize("k", range(len(golden()["cont_param"])))
def test_continuation_param(oracle_lib, k):
    e = golden()["cont_param"][k]
    spec = spec_from_hex(e["spec"])
    r = ORA.continuation_param(spec, e["step"], e["pname"], unhex(e["goal"]))
    assert (r["info"], r["solver_calls"], r["nfev_total"]) == (e["info"], e["solver_calls"], e["nfev_total"])
    assert_same(r["x"], unhex(e["x"]), "continuation " + spec["name"])


@pytest.mark.parametrize("k", range(len(golden()["cont_boundary"])))
def test_continuation_boundary(oracle_lib, k):
    e = golden()["cont_boundary"][k]
    spec = spec_from_hex(e["spec"])
    r = ORA.continuation_boundary(spec, e["step"], unhex(e["timed"]), unhex(e["Xd"]))
    assert (r["info"], r["solver_calls"], r["nfev_total"]) == (e["info"], e["solver_calls"], e["nfev_total"])
    assert_same(r["x"], unhex(e["x"]), "continuation " + spec["name"])


def test_survey_known_answers():
    """known-answer values obtained independently (reference code behind
    scipy's MINPACK) -- nfev of the demo solves and the free final times."""
    g = golden()
    want = {"di_free_tf": 82, "goddard_stage1": 1184, "goddard_stage4_singular": 101, "interceptor_init": 180}
    for e in g["solve"]:
        if e["spec"]["name"] in want:
            assert e["nfev"] == want[e["spec"]["name"]] and e["info"] == 1
    di = [e for e in g["solve"] if e["spec"]["name"] == "di_free_tf"][0]
    assert unhex(di["x"])[-1] == pytest.approx(27.655974527562908, rel=1e-13)
    g1 = [e for e in g["solve"] if e["spec"]["name"] == "goddard_stage1"][0]
    assert unhex(g1["x"])[-1] == pytest.approx(0.21965083876703931, rel=1e-14)
    assert unhex(g1["x"])[7] == pytest.approx(-2.1984225201952747, rel=1e-14)
    kd0 = unhex(g["traj"][1]["Xf"])
    assert kd0[0] == 0.99344496820998862 and kd0[13] == 0.091240570898303591
    nf = {e["spec"]["name"]: e["nfev_total"] for e in g["cont_param"]}
    assert nf["goddard_stage2_KD"] == 188 and nf["goddard_stage3_mu2"] == 638 and nf["di_cont_muT"] == 151
