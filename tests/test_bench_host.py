"""Host-side logic of bench.py that needs no GPU: the CPU pool keeps the order of its problems, and the
parity objects report what they say."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_cpu_pool_keeps_problem_order(oracle_lib):
    import bench
    import scenarios as S
    from backends import OracleBackend
    specs = []
    for k in range(5):
        s = S.di_problem()
        s["Xb"][1][1] = 15.0 + k
        specs.append(s)
    pool = bench.CpuPool(cores=3)
    try:
        _, res = pool.solve(specs)
    finally:
        pool.close()
    ora = OracleBackend()
    for s, r in zip(specs, res):
        o = ora.solve(s)
        assert (r[0], r[1]) == (o["info"], o["nfev"])
        assert np.linalg.norm(r[2] - o["x"]) <= 1e-8 * np.linalg.norm(o["x"])
    # the problems differ, so a permutation would show
    assert len({tuple(np.round(r[2], 6)) for r in res}) == 5


def test_parity_solve_counts():
    import bench
    x = np.ones((4, 3))
    ref = [(1, 10, x[0]), (1, 11, x[1] * (1 + 1e-3)), (4, 50, x[2]), (1, 12, x[3])]
    p = bench.parity_solve(np.array([1, 1, 1, 5]), np.array([10, 11, 40, 12]), x, ref, 1e-6)
    assert p["sample"] == 4
    assert p["identical_info"] == 0.5 and p["identical_info_nfev"] == 0.5
    assert p["both_converged"] == 2 and p["x_within_xtol"] == 0.5
    assert abs(p["x_rel_err_max"] - 1e-3 / (1 + 1e-3)) < 1e-9
    assert p["max_abs_nfev_diff"] == 10


def test_parity_ensemble_membership():
    import bench
    runs = [[(1, 100), (1, 120), (4, 300)], [(1, 90), (1, 90), (1, 90)]]
    ref = [(1, 100, None), (1, 90, None)]
    p = bench.parity_ensemble(np.array([4, 1]), np.array([310, 120]), ref, runs)
    assert p["reference_stable_fraction"] == 0.5
    assert p["gpu_in_ensemble"] == 0.5 and p["reference_in_ensemble"] == 1.0


def test_reference_xstar_is_the_golden_solution():
    import bench
    x = bench.reference_xstar()
    assert x.shape == (85,) and abs(x[84] - 0.21965083876703931) < 1e-15      # tf of Goddard stage 1


def test_strong_scaling_blocks_carry_every_per_problem_array():
    """bench.py --scaling strong: a rank's block of the global batch is cut out of 100000-problem blocks drawn with
    fixed seeds; every per-problem array -- the continuation targets too -- must follow the cut, whatever the
    world size (a rank whose block spans two seed blocks included)."""
    import bench
    total = 250000
    for build in (bench.wl_interceptor, bench.wl_covid19):
        full = bench._strong_block(build, None, total, 0, total, 7)
        assert full.B == total
        for world in (2, 8):
            per = -(-total // world)
            for rank in (0, world - 1, world // 2):
                lo, hi = min(rank * per, total), min(rank * per + per, total)
                w = bench._strong_block(build, None, total, lo, hi, 7)
                assert w.B == hi - lo
                assert np.array_equal(w.x0, full.x0[lo:hi]) and np.array_equal(w.Xb, full.Xb[lo:hi])
                for k in bench._PER_PROBLEM_CONT:
                    if k in w.cont:
                        assert w.cont[k].shape[0] == hi - lo, (k, w.cont[k].shape)
                        assert np.array_equal(w.cont[k], full.cont[k][lo:hi]), k
                s0 = w.spec(0)
                assert s0["cont"]["step"] == w.cont["step"]


def test_ncu_traffic_record_belongs_to_these_sources():
    """roofline.traffic is taken from profiles/r2_traffic.json only when that record was captured from the CUDA sources
    being run (digest of socp_b200/csrc): the committed record must match the committed sources, and it must carry a
    measured figure for the kernels of the Broyden and Jacobian phases."""
    import json
    import bench
    with open(bench.TRAFFIC_FILE) as f:
        t = json.load(f)
    assert t["source_digest"] == bench.source_digest(), "re-run tools/profile_r2.sh + tools/summarize_r2.py after changing csrc/"
    kernels, src = bench.ncu_traffic()
    for name in ("hybrd_chain_kernel", "hybrd_qpass_kernel", "hybrd_jac_kernel"):
        assert kernels[name]["dram_bytes_per_unit"] > 1e4 and kernels[name]["units"] > 100, name
    assert "r2_traffic.json" in src
    # the record is of the Goddard batch (P = 85): it is not applied to the kernels of another workload
    assert bench.ncu_traffic("goddard_warm", 85)[0] == kernels
    other, why = bench.ncu_traffic("covid19", 160)
    assert other == {} and "not of this one" in why
