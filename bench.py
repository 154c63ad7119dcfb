#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched shooting engine (BASELINE.json metric).

Workload (BASELINE.json configs[1], SURVEY.md section 8d "C2"): the Goddard rocket problem with free
final time (tests/testGoddard.cpp:28-99: multiple shooting M=6, P=85 unknowns, 10 RK4 steps per
segment, xtol 1e-6, KD=0, mu2=1), a batch of 1e5 perturbed initial conditions per GPU, trivial
costate guess 0.1.  One "step" = one batched solve of the whole batch (socp_solve_batch).

  python bench.py --gpus N --steps K --warmup W             (torchrun for N > 1, one rank per GPU)
  python bench.py --impl reference ...                      (the reference's CPU solver, all cores)

Prints ONE JSON line (rank 0).  `value` = shooting solves per second with the inputs resident in
HBM; `e2e` = the same through the public API with pinned HOST buffers (H2D + D2H inside the timed
region); `roofline` = the RK4 integration kernel against the FP64 FMA peak measured on this GPU;
`cpu_baseline` = the unmodified reference (oracle/_ref) timed on the host cores on a bounded
sample.  oracle/ is used here only as that CPU baseline / reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FLOPS_PER_RK4_STEP = {"goddard": 1100.0}     # SURVEY.md 8(d) / BASELINE.md section 3 (nominal count)
# DRAM traffic of hybrd_res_kernel per Broyden iteration of one problem, from the `ncu --set full` capture
# profiles/r1_final_ncu_res_raw.csv: (dram__bytes_read.sum + dram__bytes_write.sum) = 660.8 MB for the 3537
# iterations of the captured launch (1.02x the algorithmic 182 KB: no wasted re-reads)
NCU_DRAM_BYTES_PER_ITERATION = 660.8e6 / 3537
METRIC = "shooting solves/sec (Goddard free-tf, M=6, P=85; RK4 steps/sec and % of FP64 roofline alongside)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=100000, help="problems per GPU")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=0, help="problems in the CPU baseline sample (0 = 16 per core, about 10 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# workload construction (host side, untimed)
# ---------------------------------------------------------------------------------------------
def goddard_workload(eng, B, seed):
    """Config C2 arrays for B problems: (shape, mparams, time, Xb, x0).  The initial guess follows
    shooting::InitShooting (shooting.cpp:202-245): interior node states by integrating from the
    initial state with the constructor's KD=310 (tests/testGoddard.cpp sets KD=0 afterwards)."""
    import socp_b200 as sb
    import scenarios as S
    M, n, N = 6, 7, 14
    Xi, xf0 = S.goddard_batch_inputs(B, seed=seed)
    mode_t, mode_X = sb.default_modes(sb.GODDARD, M, sb.FREE, S.GODDARD_MODE_XF)
    shape = sb.make_shape(sb.GODDARD, M, mode_t, mode_X, 10)
    mp_init = np.array(S.DEFAULTS[S.GODDARD]); mp_init[6] = 1.0          # mu2 = 1, KD = 310
    ti, tf = 0.0, 0.1
    time_ = np.tile(np.array([ti + i * (tf - ti) / M for i in range(M + 1)]), (B, 1))
    x0 = np.zeros((B, N * M + 1))
    x0[:, :N] = Xi
    for i in range(1, M):
        x0[:, N * i:N * (i + 1)] = eng.traj_batch(sb.GODDARD, mp_init, ti, Xi, time_[:, i], 10)
    x0[:, N * M] = tf
    mp = np.tile(mp_init, (B, 1)); mp[:, 2] = 0.0                        # KD = 0
    Xb = np.zeros((B, M + 1, n))
    Xb[:, 0, :] = Xi[:, :n]
    Xb[:, M, 0] = xf0
    return shape, mp, time_, Xb.reshape(B, -1), x0


def spec_of(k, mp, time_, Xb, x0):
    """Problem k of the batch as a scenario spec for the CPU checkers."""
    import scenarios as S
    mode_t, mode_X = S.default_modes(S.GODDARD, 6, S.FREE, S.GODDARD_MODE_XF)
    return S.make_spec(S.GODDARD, 6, mode_t, mode_X, time_[k], Xb[k].reshape(7, 7), x0[k], 1e-6,
                       mparams=mp[k], steps=10, name="goddard_batch_%d" % k)


# ---------------------------------------------------------------------------------------------
# CPU reference (oracle/_ref = unmodified reference sources; falls back to the C port)
# ---------------------------------------------------------------------------------------------
def _cpu_worker(specs):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import pyref
    import backends
    be = backends.RefBackend() if pyref.available() else backends.OracleBackend()
    out = []
    for s in specs:
        r = be.solve(s)
        out.append((int(r["info"]), int(r["nfev"])))
    return out


class CpuPool:
    """One process per host core, each with its own copy of the reference library."""

    def __init__(self, cores=None):
        import multiprocessing as mp
        self.cores = cores or len(os.sched_getaffinity(0))
        self.pool = mp.get_context("spawn").Pool(self.cores)
        from oracle import pyref
        self.kind = "reference" if pyref.available() else "port"

    def solve(self, specs):
        chunks = [specs[i::self.cores] for i in range(self.cores)]
        chunks = [c for c in chunks if c]
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, chunks)
        dt = time.perf_counter() - t0
        flat = [r for c in res for r in c]
        return dt, flat

    def close(self):
        self.pool.close()
        self.pool.join()


# ---------------------------------------------------------------------------------------------
def hbm_peak_gbs():
    """Measured copy bandwidth of this pool's B200s (driver-written MEASURED_PEAKS.json), else the
    fallback stated in B200_PROFILING.md."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[3 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]),
                "power_w_max": max(float(s[2]) for s in self.samples), "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def run_reference(args, rank):
    """The reference's own CPU implementation (oracle/_ref) on the host cores, same workload."""
    if rank != 0:
        return
    import scenarios as S
    from backends import OracleBackend
    pool = CpuPool()
    per_step = args.cpu_sample or 8 * pool.cores
    n = per_step * (args.steps + args.warmup)
    # build the same problems as the GPU arm (same seed); the guess integration runs on the CPU port
    ora = OracleBackend()
    Xi, xf0 = S.goddard_batch_inputs(args.batch, seed=20260002)
    specs = []
    for k in range(n):
        kk = k % args.batch
        specs.append(S.goddard_problem(lambda mp, a, b, c: ora.traj(S.GODDARD, mp, a, b, c, 10), Xi=Xi[kk], xf0=xf0[kk]))
    for w in range(args.warmup):
        pool.solve(specs[w * per_step:(w + 1) * per_step])
    t_total, res = 0.0, []
    for s in range(args.steps):
        lo = (args.warmup + s) * per_step
        dt, r = pool.solve(specs[lo:lo + per_step])
        t_total += dt
        res += r
    pool.close()
    value = per_step * args.steps / t_total
    nfev = sum(r[1] for r in res)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "goddard_free_tf_M6_P85_batch (configs[1])", "batch_per_gpu": args.batch,
                   "sample_per_step": per_step},
        "rk4_steps_per_s": nfev * 60.0 / t_total,
        "converged_fraction": sum(1 for r in res if r[0] == 1) / len(res),
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": pool.cores, "kind": pool.kind,
                         "sample": "%d problems of the batch per step, one process per core" % per_step},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    # stdout carries exactly ONE JSON line: anything libraries print there (NCCL's version banner, ...) goes
    # to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import socp_b200 as sb
    from socp_b200 import sharding
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    # CPU baseline pool first (spawned before the heavy CUDA work; rank 0, N = 1 only)
    pool = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        pool = CpuPool()

    eng = sb.Engine(local)
    peak_gflops, clk = eng.measure_fp64_peak()
    B = args.batch
    shape, mp, time_, Xb, x0 = goddard_workload(eng, B, seed=20260002 + rank + int(os.environ.get("SOCP_BENCH_SEED_OFFSET", "0")))
    P = x0.shape[1]

    def dv(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_mp, d_time, d_Xb, d_x0 = dv(mp), dv(time_), dv(Xb), dv(x0)
    d_x = torch.empty_like(d_x0)
    d_info = torch.empty(B, dtype=torch.int32, device=dev)
    d_nfev = torch.empty(B, dtype=torch.int32, device=dev)
    d_fnorm = torch.empty(B, dtype=torch.float64, device=dev)
    eng.use_torch_stream()

    def step():
        d_x.copy_(d_x0)
        eng.solve_batch(shape, d_mp, d_time, d_Xb, d_x, xtol=1e-6, maxfev=10000, info=d_info, nfev=d_nfev, fnorm=d_fnorm)
        if world > 1:
            # the only exchange of the path: gather converged unknowns + status over NVLink
            sharding.gather_results(d_x, d_info, d_nfev)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    eng.reset_stats()
    eng.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = eng.stats()
    eng.set_profiling(False)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if world > 1:
        print("rank %d: %.1f ms for %d step(s), %d solver rounds, res %.0f jac %.0f ms" %
              (rank, ms, args.steps, st["solver_rounds"], st["advance_ms"] - st["jac_ms"], st["jac_ms"]), file=sys.stderr)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    agg = torch.tensor([st["rk4_steps"], float((d_info == 1).sum().item()), float(d_nfev.sum().item()),
                        float(((d_info == 1) & (d_fnorm < 1e-5)).sum().item())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    rk4_steps_per_step = float(agg[0].item()) / args.steps
    rk4_rate = rk4_steps_per_step / (ms_per_step * 1e-3)

    # rooflines, per kernel of the solver round (CUDA events around every launch on the engine stream)
    flops = FLOPS_PER_RK4_STEP["goddard"]
    peak_tf = peak_gflops / 1e3
    hbm_peak, hbm_src = hbm_peak_gbs()
    int_ms, int_n = st["integrate_ms"], max(st["integrate_launches"], 1.0)
    jac_ms, res_ms, asm_ms = st["jac_ms"], st["advance_ms"] - st["jac_ms"], st["assemble_ms"]
    LR = P * (P + 1) // 2
    # hybrd_res_kernel, per Broyden iteration of one problem: Q read + written once (2 P^2), packed R
    # read + written once (2 LR), seven work vectors in and five out (12 P)   [DESIGN.md section 5]
    bytes_iter = 8.0 * (2 * P * P + 2 * LR + 12 * P)
    # hybrd_jac_kernel, per factorisation: Householder QR (4/3 P^3) + accumulation of Q (4/3 P^3)
    flops_jac = 8.0 / 3.0 * P ** 3
    kern = {
        "integrate_worklist<goddard>": {"bound": "fp64", "ms": int_ms, "unit": "TFLOP/s", "peak": peak_tf,
                                        "achieved": st["rk4_steps"] * flops / (int_ms * 1e-3) / 1e12 if int_ms > 0 else 0.0},
        "hybrd_res_kernel": {"bound": "hbm", "ms": res_ms, "unit": "GB/s", "peak": hbm_peak,
                             "achieved": st["iterations"] * bytes_iter / (res_ms * 1e-3) / 1e9 if res_ms > 0 else 0.0},
        "hybrd_jac_kernel": {"bound": "fp64", "ms": jac_ms, "unit": "TFLOP/s", "peak": peak_tf,
                             "achieved": st["jac_evals"] * flops_jac / (jac_ms * 1e-3) / 1e12 if jac_ms > 0 else 0.0},
        "assemble_kernel<goddard>": {"bound": "latency", "ms": asm_ms, "unit": None, "peak": None, "achieved": None},
    }
    for k in kern.values():
        k["share_of_step"] = k["ms"] / ms if ms > 0 else None
        k["frac"] = (k["achieved"] / k["peak"]) if k["peak"] else None
    dom = max((k for k in kern if kern[k]["peak"]), key=lambda k: kern[k]["ms"])
    d = kern[dom]
    n_launch = max(st["solver_rounds"], 1.0)
    roofline = {"kernel": dom, "bound": d["bound"], "achieved": d["achieved"], "peak": d["peak"], "unit": d["unit"],
                "frac": d["frac"],
                "traffic": (NCU_DRAM_BYTES_PER_ITERATION * st["iterations"] / n_launch) if dom == "hybrd_res_kernel" else None,
                "traffic_note": "bytes per launch = ncu-measured DRAM bytes per problem-iteration (profiles/r1_final_ncu_res_raw.csv) "
                                "x problem-iterations per launch of this run",
                "algorithmic_bytes_per_launch": bytes_iter * st["iterations"] / n_launch,
                "avg_launch_ms": d["ms"] / n_launch, "launches": n_launch,
                "share_of_step": d["share_of_step"],
                "algorithmic_bytes_per_problem_iteration": bytes_iter,
                "problem_iterations": st["iterations"], "jacobian_factorisations": st["jac_evals"],
                "peak_source": {"hbm": hbm_src, "fp64": "measured on this GPU: register-resident DFMA chain "
                                "(socp_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry"},
                "flops_per_rk4_step": flops, "kernels": kern}

    # the RK4 trajectory kernel alone (BASELINE metric "RK4 steps/sec, % of FP64 roofline"): the Goddard
    # batch of SURVEY 8d's microbenchmark, 2^20 trajectories of one 10-step segment, device resident
    rk4_kernel = None
    if rank == 0 or world > 1:
        Bt = 1 << 20
        reps = 6
        Xi_t = torch.from_numpy(np.tile(x0[:1, :14], (Bt, 1)) * (1 + 1e-3 * np.random.default_rng(7).uniform(-1, 1, (Bt, 14)))).to(dev)
        mp_t = d_mp[:1].repeat(Bt, 1).contiguous()
        t0_t = torch.zeros(Bt, dtype=torch.float64, device=dev)
        tf_t = torch.full((Bt,), 0.1 / 6, dtype=torch.float64, device=dev)
        out_t = torch.empty_like(Xi_t)
        for _ in range(3):
            eng.traj_batch(sb.GODDARD, mp_t, t0_t, Xi_t, tf_t, 10, out=out_t)
        eng.sync()
        eng.reset_stats()
        eng.timer_start()
        for _ in range(reps):
            eng.traj_batch(sb.GODDARD, mp_t, t0_t, Xi_t, tf_t, 10, out=out_t)
        t_ms = eng.timer_stop()
        steps_t = eng.stats()["rk4_steps"]
        tf_ach = steps_t * flops / (t_ms * 1e-3) / 1e12
        rk4_kernel = {"kernel": "traj_kernel<goddard>", "trajectories": Bt, "rk4_steps_per_s": steps_t / (t_ms * 1e-3),
                      "ms_per_launch": t_ms / reps, "achieved": tf_ach, "peak": peak_tf, "unit": "TFLOP/s",
                      "frac": tf_ach / peak_tf, "bound": "fp64", "flops_per_rk4_step": flops,
                      "note": "inputs 117 MB + outputs 117 MB per launch (> L2 together with the solver workspace)"}
        del Xi_t, mp_t, t0_t, tf_t, out_t

    # end to end through the public API with pinned host buffers
    e2e = None
    if rank == 0 or world > 1:
        h = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in
             dict(mp=mp, time=time_, Xb=Xb, x0=x0).items()}
        hx = torch.empty_like(h["x0"]).pin_memory()
        h_info = torch.empty(B, dtype=torch.int32).pin_memory()
        h_nfev = torch.empty(B, dtype=torch.int32).pin_memory()
        h_fn = torch.empty(B, dtype=torch.float64).pin_memory()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            hx.copy_(h["x0"])
            eng.solve_batch(shape, h["mp"].numpy(), h["time"].numpy(), h["Xb"].numpy(), hx.numpy(), xtol=1e-6,
                            maxfev=10000, info=h_info.numpy(), nfev=h_nfev.numpy(), fnorm=h_fn.numpy())
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d = (mp.nbytes + time_.nbytes + Xb.nbytes + x0.nbytes)
        d2h = (x0.nbytes + 4 * B + 4 * B + 8 * B)
        e2e = {"value": world * B * args.e2e_steps / dt, "unit": "solves/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
               "api": "socp_solve_batch(mem=SOCP_HOST) via socp_b200.Engine.solve_batch, pinned host buffers"}

    cpu_baseline = None
    if pool is not None:
        n_s = args.cpu_sample or 16 * pool.cores
        specs = [spec_of(k, mp, time_, Xb, x0) for k in range(n_s)]
        dt, res = pool.solve(specs)
        pool.close()
        nf = sum(r[1] for r in res)
        # parity spot check on the same sample (info codes as the reference reports them)
        g_info = d_info[:n_s].cpu().numpy()
        same = sum(1 for k in range(n_s) if res[k][0] == g_info[k])
        cpu_baseline = {"value": n_s / dt, "unit": "solves/s", "cores": pool.cores, "kind": pool.kind,
                        "sample": "first %d problems of the same batch, one process per core, wall time" % n_s,
                        "rk4_steps_per_s": nf * 60.0 / dt,
                        "converged_fraction": sum(1 for r in res if r[0] == 1) / n_s,
                        "same_info_as_gpu": "%d/%d" % (same, n_s)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "goddard_free_tf_M6_P85_batch (BASELINE configs[1], SURVEY C2)",
                       "batch_per_gpu": B, "global_batch": world * B, "xtol": 1e-6, "rk4_steps_per_segment": 10,
                       "parallelism": "problems sharded, %d rank(s), no data-path collective; results all-gathered" % world,
                       "l2": "working set %.1f GB per GPU >> 126 MB L2, no flush needed" % (st["device_bytes"] / 1e9)},
            "rk4_steps_per_s": rk4_rate, "rk4_steps_per_step": rk4_steps_per_step,
            "converged_fraction": float(agg[1].item()) / (world * B),            # info == 1, what SOCP accepts
            "true_root_fraction": float(agg[3].item()) / (world * B),            # info == 1 and |F| < 1e-5
            "mean_nfev": float(agg[2].item()) / (world * B),
            "solver_rounds_per_step": st["solver_rounds"] / args.steps,
            "gpu_launches": int(st["kernel_launches"]),
            "roofline": roofline, "rk4_kernel": rk4_kernel, "e2e": e2e, "cpu_baseline": cpu_baseline,
            "clocks": sampler.summary(), "fp64_peak_probe_sm_mhz": clk,
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
