// socp_b200/csrc/solver_api.inl -- host side of the batched residual / Jacobian / solve entry
// points (included at the end of api.cu).  Host glue only: shapes, work-item tables, workspace
// carving, the round loop and the continuation state machines; no arithmetic of the hot path.

namespace {

struct Carver {
    char *base;
    size_t off = 0;
    explicit Carver(void *b) : base((char *)b) {}
    template <typename T> T *take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T *p = base ? (T *)(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

struct HostPlan {
    SolverDev D;
    std::vector<int> jac_col, jac_seg, col_item0, col_nseg;
    size_t bytes_per_problem = 0;
    size_t table_ints = 0;
};

int make_plan(socp_ctx *ctx, const socp_shape *shape, HostPlan &pl) {
    if (!shape) return fail(ctx, SOCP_ERR_ARG, "shape is NULL");
    const int P = socp_num_param(shape);
    if (P < 0) return fail(ctx, SOCP_ERR_ARG, "bad shape (model id / numMulti)");
    SolverDev &D = pl.D;
    memset(&D, 0, sizeof D);
    D.model_id = shape->model_id;
    D.dim = kDim[D.model_id];
    D.N = 2 * D.dim;
    D.M = shape->num_multi;
    D.S = shape->step_nbr > 0 ? shape->step_nbr : kSteps[D.model_id];
    D.P = P;
    D.np = kNP[D.model_id];
    D.REC = D.N + 2;
    D.LR = P * (P + 1) / 2;
    D.LRS = (D.LR + 1) & ~1;
    D.QS = (P * P + 1) & ~1;
    D.nfree = P - D.N * D.M;
    D.ode_tol = (shape->integrator == SOCP_DOPRI5) ? shape->ode_tol : 0.;
    if (shape->integrator != SOCP_RK4 && shape->integrator != SOCP_DOPRI5) return fail(ctx, SOCP_ERR_ARG, "bad shape.integrator");
    if (shape->integrator == SOCP_DOPRI5 && !(shape->ode_tol > 0.)) return fail(ctx, SOCP_ERR_ARG, "shape.ode_tol must be > 0 with SOCP_DOPRI5");
    for (int j = 0; j <= D.M; ++j) {
        const int mt = shape->mode_t[j];
        if (mt < SOCP_FIXED || mt > SOCP_CONTINUOUS) return fail(ctx, SOCP_ERR_ARG, "bad mode_t");
        D.mode_t[j] = mt;
        for (int k = 0; k < D.dim; ++k) D.mode_X[j][k] = shape->mode_X[j][k];
    }
    if (D.mode_t[0] == SOCP_CONTINUOUS || D.mode_t[D.M] == SOCP_CONTINUOUS)
        return fail(ctx, SOCP_ERR_ARG, "the first and last node times must be FIXED or FREE");
    // Jacobian work items: a state/costate unknown of node i only moves segment i; a free time
    // moves the whole time line, so every segment is re-integrated for it.
    pl.col_item0.assign(P, 0);
    pl.col_nseg.assign(P, 1);
    for (int j = 0; j < P; ++j) {
        pl.col_item0[j] = (int)pl.jac_col.size();
        if (j < D.N * D.M) {
            pl.jac_col.push_back(j);
            pl.jac_seg.push_back(j / D.N);
        } else {
            pl.col_nseg[j] = D.M;
            for (int s = 0; s < D.M; ++s) { pl.jac_col.push_back(j); pl.jac_seg.push_back(s); }
        }
    }
    D.nJ = (int)pl.jac_col.size();
    pl.table_ints = 2 * (size_t)D.nJ + 2 * (size_t)P;
    // persistent bytes per problem
    size_t dbl = (size_t)P * 10 + 4 * (size_t)P + (size_t)D.QS + D.LRS + (size_t)(2 * D.M + D.nJ) * D.REC + D_COUNT;
    pl.bytes_per_problem = dbl * sizeof(double) + (I_COUNT + 4) * sizeof(int) + 2048 / 64;
    return SOCP_OK;
}

// carve the solver workspace for a wave of Bw problems; returns the total size
size_t carve(HostPlan &pl, void *blob, long Bw) {
    SolverDev &D = pl.D;
    Carver c(blob);
    const size_t P = D.P;
    int *tab = c.take<int>(pl.table_ints);
    D.jac_col = tab;
    D.jac_seg = tab ? tab + D.nJ : nullptr;
    D.col_item0 = tab ? tab + 2 * D.nJ : nullptr;
    D.col_nseg = tab ? tab + 2 * D.nJ + P : nullptr;
    D.x = c.take<double>(Bw * P); D.xe = c.take<double>(Bw * P); D.fvec = c.take<double>(Bw * P);
    D.diag = c.take<double>(Bw * P); D.qtf = c.take<double>(Bw * P);
    D.wa1 = c.take<double>(Bw * P); D.wa2 = c.take<double>(Bw * P); D.wa3 = c.take<double>(Bw * P);
    D.wa4 = c.take<double>(Bw * P); D.scr = c.take<double>(Bw * 4 * P);
    D.fjac = c.take<double>(Bw * (size_t)D.QS);
    D.r = c.take<double>(Bw * (size_t)D.LRS);
    D.ends = c.take<double>(Bw * 2 * (size_t)D.M * D.REC);
    D.jends = c.take<double>(Bw * (size_t)D.nJ * D.REC);
    D.dstate = c.take<double>(Bw * D_COUNT);
    D.istate = c.take<int>(Bw * I_COUNT);
    D.lists = c.take<int>(4 * (size_t)Bw);
    D.counts = c.take<int>(8);
    D.self = c.take<SolverDev>(1);
    return c.off + 256;
}

const int kProfSlots = 16;      // rounds between two harvests of the profiling events
const int kProfEv = 6;          // events per profiled round: start, integrate, assemble, Q pass, Broyden, Jacobian

cudaEvent_t prof_event(socp_ctx *ctx, int idx) {
    while ((int)ctx->prof_events.size() <= idx) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ctx->prof_events.push_back(e);
    }
    return ctx->prof_events[idx];
}

// accumulate the per-kernel times of the last `n` profiled rounds (events must have completed)
void prof_harvest(socp_ctx *ctx, int n, FILE *log = nullptr, long round0 = 0, const int *counts = nullptr) {
    for (int k = 0; k < n; ++k) {
        float a = 0, b = 0, q = 0, c = 0, d = 0;
        cudaEvent_t *e = &ctx->prof_events[kProfEv * k];
        cudaEventElapsedTime(&a, e[0], e[1]);          // integrate_worklist (+ variational kernels)
        cudaEventElapsedTime(&b, e[1], e[2]);          // zero_fjac + assemble
        cudaEventElapsedTime(&q, e[2], e[3]);          // Q pass (split build; ~0 otherwise)
        cudaEventElapsedTime(&c, e[2], e[4]);          // Broyden phase: Q pass + chain kernel, or the fused kernel
        cudaEventElapsedTime(&d, e[4], e[5]);          // Jacobian phase
        ctx->integrate_ms += a; ctx->integrate_launches += 1;
        ctx->assemble_ms += b;
        ctx->qpass_ms += q;
        ctx->advance_ms += c + d; ctx->advance_launches += 2;
        ctx->jac_ms += d;
        if (log) fprintf(log, "%ld %d %d %.4f %.4f %.4f %.4f %.4f\n", round0 + k, counts ? counts[0] : -1, counts ? counts[1] : -1, a, b, q, c - q, d);
    }
}

// shared-memory plan of the Powell-hybrid kernels for this problem size
struct SmemPlan {
    int G;                               // threads per problem: one warp, or a 128-thread CTA
    int groups;                          // problems per CTA (G == 32 only)
    bool stage_r, stage_q_jac;
    bool jac_r_global;                   // Jacobian phase: leave R in global memory when that fits one more CTA per SM
    int doubles_res, doubles_jac;        // shared doubles per group
    size_t bytes_res, bytes_jac;         // per CTA
    // split Broyden phase (hybrd_qpass_kernel + hybrd_chain_kernel): Q fits one CTA's shared memory
    bool split;
    size_t bytes_qpass;                  // per CTA of the Q pass
    int chain_doubles;                   // per problem (= per warp) of the chain kernel
    int qpass_per_sm, chain_per_sm;      // resident CTAs / warps per SM
};

// One warp per problem up to P = 32, a 128-thread CTA above (measured at P = 85: one warp per problem
// with R left in global memory is within 13% of the CTA variant, a register-resident Householder QR
// was 1.7x slower than the shared-memory one -- see DESIGN.md section 7).  R and the work vectors
// live in shared memory when they fit, Q too in the Jacobian phase.
SmemPlan smem_plan(const SolverDev &D) {
    const size_t limit = 225 * 1024;
    const size_t base = 8 + 13 * (size_t)D.P;
    const size_t ldq = (size_t)(D.P | 1);
    SmemPlan p;
    p.G = (D.P <= 32) ? 32 : 128;
    p.stage_r = (base + D.LR) * 8 <= limit;
    const size_t dr = base + (p.stage_r ? D.LR : 0);
    // Jacobian phase with Q staged: [vectors][R][pad to 16 B][Q, ld = P | 1][8 P reflector buffers of qrfac_w / qform_w]
    const size_t qs_jac = 1 + ((ldq * D.P + 1) & ~(size_t)1) + 8 * (size_t)D.P;
    p.stage_q_jac = p.stage_r && (dr + qs_jac) * 8 <= limit;
    const size_t dj = dr + (p.stage_q_jac ? qs_jac : 0);
    p.doubles_res = (int)dr;
    p.doubles_jac = (int)dj;
    p.groups = 1;
    if (p.G == 32) {
        // small problems: pack up to 4 warps (problems) in a CTA while shared memory allows
        while (p.groups < 4 && dj * 8 * (p.groups * 2) <= 48 * 1024) p.groups *= 2;
    }
    p.bytes_res = dr * 8 * p.groups;
    p.bytes_jac = dj * 8 * p.groups;
    // In the Jacobian phase R is written once (pack_r) and read by one dogleg, while Q is swept 2 P times:
    // if dropping R from shared memory lets one more CTA live on the SM, do that (P = 85: 96 -> 67 KB,
    // 2 -> 3 CTAs/SM, measured -25 % kernel time)
    const size_t sm_bytes = 227 * 1024;
    p.jac_r_global = p.G == 128 && p.stage_q_jac &&
                     sm_bytes / ((dj - D.LR) * 8 + 1024) > sm_bytes / (dj * 8 + 1024);
    // Split Broyden phase for 32 < P: the Q pass stages the whole Q (QS doubles) + F(x+p) + 4 P coefficients
    // + the mbarrier in one CTA; the chain kernel needs R + 13 vectors per warp.  SOCP_BROYDEN=fused keeps the
    // one-kernel form (A/B measurements, and the fallback when Q does not fit: P > ~165).
    p.bytes_qpass = ((size_t)D.QS + ((D.P + 1) & ~1) + 4 * (size_t)D.P + 2) * 8;
    // chain kernel, lean layout: [8][9 P vectors][pad to 16 B][R, LRS doubles]
    p.chain_doubles = (int)(((8 + 9 * (size_t)D.P + 1) & ~(size_t)1) + D.LRS);
    const char *mode = getenv("SOCP_BROYDEN");
    p.split = p.G == 128 && p.stage_r && p.bytes_qpass <= limit && !(mode && !strcmp(mode, "fused"));
    p.qpass_per_sm = (int)std::max<size_t>(1, sm_bytes / (p.bytes_qpass + 1024));
    p.chain_per_sm = (int)std::min<size_t>(16, std::max<size_t>(1, sm_bytes / ((size_t)p.chain_doubles * 8 + 1024)));
    return p;
}

// Opt-in dynamic shared memory is a per-DEVICE function attribute: the size already configured is kept per
// context (one context = one device), never process-wide, so a second context on another GPU of the same
// process configures its own copy and two host threads never share the table.
template <typename K>
int launch_smem(socp_ctx *ctx, K kernel, int grid, int threads, size_t smem, const SolverDev &D, int cur, int doubles,
                cudaStream_t st = nullptr) {
    if (smem > 48 * 1024) {
        size_t &have = ctx->smem_configured[(const void *)kernel];
        if (smem > have) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) {
                ctx->err = std::string("cudaFuncSetAttribute(MaxDynamicSharedMemorySize): ") + cudaGetErrorString(e);
                ctx->launch_error = SOCP_ERR_CUDA;
                return SOCP_ERR_CUDA;
            }
            have = smem;
        }
    }
    kernel<<<grid, threads, smem, st ? st : ctx->stream>>>(D, cur, doubles);
    return SOCP_OK;
}

int launch_qpass(socp_ctx *ctx, int grid, size_t smem, const SolverDev &D, int cur) {
    if (smem > 48 * 1024) {
        size_t &have = ctx->smem_configured[(const void *)hybrd_qpass_kernel];
        if (smem > have) {
            cudaError_t e = cudaFuncSetAttribute(hybrd_qpass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) {
                ctx->err = std::string("cudaFuncSetAttribute(hybrd_qpass_kernel): ") + cudaGetErrorString(e);
                ctx->launch_error = SOCP_ERR_CUDA;
                return SOCP_ERR_CUDA;
            }
            have = smem;
        }
    }
    hybrd_qpass_kernel<<<grid, 128, smem, ctx->stream>>>(D, cur);
    return SOCP_OK;
}

void launch_hybrd(socp_ctx *ctx, const SolverDev &D, int cur, int grid, int prof_slot) {
    const SmemPlan sp = smem_plan(D);
    const int thr = (sp.G == 32) ? 32 * sp.groups : 128;
    const int g = (sp.G == 32) ? grid * (4 / sp.groups) : grid;
    if (sp.G == 32) {
        if (prof_slot >= 0) cudaEventRecord(prof_event(ctx, kProfEv * prof_slot + 3), ctx->stream);
        if (sp.stage_r) launch_smem(ctx, hybrd_res_kernel<32, true>, g, thr, sp.bytes_res, D, cur, sp.doubles_res);
        else launch_smem(ctx, hybrd_res_kernel<32, false>, g, thr, sp.bytes_res, D, cur, sp.doubles_res);
        if (prof_slot >= 0) cudaEventRecord(prof_event(ctx, kProfEv * prof_slot + 4), ctx->stream);
        if (sp.stage_q_jac) launch_smem(ctx, hybrd_jac_kernel<32, true, true>, g, thr, sp.bytes_jac, D, cur, sp.doubles_jac);
        else if (sp.stage_r) launch_smem(ctx, hybrd_jac_kernel<32, true, false>, g, thr, sp.bytes_jac, D, cur, sp.doubles_jac);
        else launch_smem(ctx, hybrd_jac_kernel<32, false, false>, g, thr, sp.bytes_jac, D, cur, sp.doubles_jac);
    } else {
        const bool full = g == D.sm_count * 6;
        // The Broyden phase and the Jacobian phase of a round work on disjoint problems: outside profiled passes the
        // Jacobian kernel runs on a second stream, forked after the assembly and joined before the next round, so that
        // the partial waves of the two latency-bound kernels fill each other's idle SMs.  (SOCP_OVERLAP=0: one stream.)
        static const bool overlap_on = !(getenv("SOCP_OVERLAP") && !strcmp(getenv("SOCP_OVERLAP"), "0"));
        const bool overlap = overlap_on && prof_slot < 0 && ctx->stream2 != nullptr;
        cudaStream_t sj = overlap ? ctx->stream2 : ctx->stream;
        auto launch_jac = [&]() {
            if (sp.stage_q_jac && sp.jac_r_global)          // R straight to global memory: one more CTA per SM
                launch_smem(ctx, hybrd_jac_kernel<128, false, true>, g, thr, (size_t)(sp.doubles_jac - D.LR) * 8, D, cur, sp.doubles_jac - D.LR, sj);
            else if (sp.stage_q_jac) launch_smem(ctx, hybrd_jac_kernel<128, true, true>, g, thr, sp.bytes_jac, D, cur, sp.doubles_jac, sj);
            else if (sp.stage_r) launch_smem(ctx, hybrd_jac_kernel<128, true, false>, g, thr, sp.bytes_jac, D, cur, sp.doubles_jac, sj);
            else launch_smem(ctx, hybrd_jac_kernel<128, false, false>, g, thr, sp.bytes_jac, D, cur, sp.doubles_jac, sj);
        };
        if (overlap) {
            cudaEventRecord(ctx->ev_fork, ctx->stream);
            cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0);
            launch_jac();
            cudaEventRecord(ctx->ev_join, ctx->stream2);
        }
        if (sp.split) {
            // Q pass: one CTA per problem, a whole number of waves of resident CTAs when the round is full
            const int gq = full ? D.sm_count * sp.qpass_per_sm * 2 : g;
            launch_qpass(ctx, gq, sp.bytes_qpass, D, cur);
            if (prof_slot >= 0) cudaEventRecord(prof_event(ctx, kProfEv * prof_slot + 3), ctx->stream);
            // chain kernel: one warp (= one 32-thread CTA) per problem
            // SOCP_CHAIN_R=global: the packed factor stays in global memory (L2), 16 problems per SM instead of 5
            static const bool r_global = getenv("SOCP_CHAIN_R") && !strcmp(getenv("SOCP_CHAIN_R"), "global");
            if (r_global) {
                const int cd = 8 + 13 * D.P;
                const int gc = full ? D.sm_count * 16 * 2 : g;
                launch_smem(ctx, hybrd_chain_kernel<false>, gc, 32, (size_t)cd * 8, D, cur, cd);
            } else {
                const int gc = full ? D.sm_count * sp.chain_per_sm * 2 : g;
                launch_smem(ctx, hybrd_chain_kernel<true>, gc, 32, (size_t)sp.chain_doubles * 8, D, cur, sp.chain_doubles);
            }
            ctx->launches += 1;
        } else {
        if (prof_slot >= 0) cudaEventRecord(prof_event(ctx, kProfEv * prof_slot + 3), ctx->stream);
        // the round grid is 6 CTAs per SM (two waves of the 3-CTA/SM Jacobian phase); the Broyden phase holds
        // 4 CTAs/SM (P = 85), so a full grid becomes 8 per SM: two full waves again
        const int g_res = full ? D.sm_count * 8 : g;
        if (sp.stage_r) launch_smem(ctx, hybrd_res_kernel<128, true>, g_res, thr, sp.bytes_res, D, cur, sp.doubles_res);
        else launch_smem(ctx, hybrd_res_kernel<128, false>, g_res, thr, sp.bytes_res, D, cur, sp.doubles_res);
        }
        if (prof_slot >= 0) cudaEventRecord(prof_event(ctx, kProfEv * prof_slot + 4), ctx->stream);
        if (overlap) cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
        else launch_jac();
    }
}

// hybrj: segment sensitivities and the analytic Jacobian for the problems in the Jacobian list
template <int MODEL>
typename std::enable_if<Variational<MODEL>::HAS>::type launch_variational(socp_ctx *ctx, const SolverDev &D, int cur, int grid) {
    integrate_var_worklist<MODEL><<<grid, 128, 0, ctx->stream>>>(D, cur);
    assemble_jac_kernel<MODEL><<<grid, 128, 0, ctx->stream>>>(D, cur);
    ctx->launches += 2;
}
template <int MODEL>
typename std::enable_if<!Variational<MODEL>::HAS>::type launch_variational(socp_ctx *, const SolverDev &, int, int) {}

template <int MODEL>
void launch_round(socp_ctx *ctx, const SolverDev &D, int cur, int grid_int, int grid_adv, int prof_slot, long coop_items) {
    if (prof_slot >= 0) cudaEventRecord(prof_event(ctx, kProfEv * prof_slot), ctx->stream);
    constexpr int L = Coop<MODEL>::LANES;
    if (coop_items >= 0 && use_coop(ctx, coop_items, L)) {
        // few work items (the tail of a solve, a small batch): a cooperative group of L lanes per segment
        const int gc = (int)std::max<long>(1, std::min<long>(grid_int * (long)L, (coop_items * L + 127) / 128));
        if (D.ode_tol > 0.) integrate_worklist<MODEL, true, L><<<gc, 128, 0, ctx->stream>>>(D, cur);
        else integrate_worklist<MODEL, false, L><<<gc, 128, 0, ctx->stream>>>(D, cur);
    } else if (D.ode_tol > 0.) integrate_worklist<MODEL, true><<<grid_int, 128, 0, ctx->stream>>>(D, cur);
    else integrate_worklist<MODEL, false><<<grid_int, 128, 0, ctx->stream>>>(D, cur);
    if (D.analytic) launch_variational<MODEL>(ctx, D, cur, grid_int);
    if (prof_slot >= 0) cudaEventRecord(prof_event(ctx, kProfEv * prof_slot + 1), ctx->stream);
    if (!D.analytic) { zero_fjac_kernel<<<grid_int, 256, 0, ctx->stream>>>(D, cur); ctx->launches += 1; }
    assemble_kernel<MODEL><<<grid_int, 128, 0, ctx->stream>>>(D, cur);
    if (prof_slot >= 0) cudaEventRecord(prof_event(ctx, kProfEv * prof_slot + 2), ctx->stream);
    launch_hybrd(ctx, D, cur, grid_adv, prof_slot);
    if (prof_slot >= 0) cudaEventRecord(prof_event(ctx, kProfEv * prof_slot + 5), ctx->stream);
    ctx->launches += 4;
    ctx->rounds += 1;
}

void launch_round_any(socp_ctx *ctx, const SolverDev &D, int cur, int gi, int ga, int ps, long coop_items) {
    switch (D.model_id) {
    case SOCP_GODDARD: launch_round<GODDARD>(ctx, D, cur, gi, ga, ps, coop_items); break;
    case SOCP_DOUBLE_INTEGRATOR: launch_round<DOUBLE_INTEGRATOR>(ctx, D, cur, gi, ga, ps, coop_items); break;
    case SOCP_COVID19: launch_round<COVID19>(ctx, D, cur, gi, ga, ps, coop_items); break;
    case SOCP_VTOL_UAV: launch_round<VTOL_UAV>(ctx, D, cur, gi, ga, ps, coop_items); break;
    case SOCP_INTERCEPTOR: launch_round<INTERCEPTOR>(ctx, D, cur, gi, ga, ps, coop_items); break;
    }
}

// Run the state machine for B problems given DEVICE pointers; waves sized to the free memory.
int run_solver(socp_ctx *ctx, const socp_shape *shape, long B, const double *d_mparams, const double *d_time,
               const double *d_Xb, const double *d_x_in, int run_mode, double xtol, int maxfev, double epsfcn,
               double *d_x_out, double *d_fvec_out, double *d_fjac_out, int *d_info, int *d_nfev, double *d_fnorm,
               int analytic = 0, int *d_njev = nullptr, const int *d_active = nullptr, long live_hint = -1) {
    HostPlan pl;
    int rc = make_plan(ctx, shape, pl);
    if (rc != SOCP_OK) return rc;
    if (analytic && shape->model_id != SOCP_DOUBLE_INTEGRATOR)
        return fail(ctx, SOCP_ERR_UNSUPPORTED, "analytic Jacobian (modelOrder 1): only the double integrator has variational equations, as in the reference");
    pl.D.analytic = analytic;
    if (B == 0) return SOCP_OK;
    SolverDev &D = pl.D;
    // wave size from the memory that is free now plus what the workspace already holds
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(ctx, cudaMemGetInfo(&free_b, &total_b));
    const size_t budget = (size_t)((free_b + ctx->solver.cap) * 0.85);
    const size_t reserve = (size_t)1 << 20;
    if (budget < reserve + pl.bytes_per_problem)
        return fail(ctx, SOCP_ERR_NOMEM, "not enough device memory for one problem (" + std::to_string(pl.bytes_per_problem) +
                                         " bytes per problem, " + std::to_string(free_b) + " free)");
    const long wave = (long)std::min<size_t>((size_t)B, (budget - reserve) / pl.bytes_per_problem);
    size_t need = carve(pl, nullptr, wave);
    if (need > ctx->solver.cap) {
        if (ctx->solver.blob) cudaFree(ctx->solver.blob);
        ctx->solver.blob = nullptr;
        ctx->solver.cap = 0;
        if (cudaMalloc(&ctx->solver.blob, need) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, SOCP_ERR_NOMEM, "cudaMalloc failed for the solver workspace (" + std::to_string(need) + " bytes)");
        }
        ctx->solver.cap = need;
    }
    if (!ctx->solver.h_counts) {
        CUDA_TRY(ctx, cudaHostAlloc((void **)&ctx->solver.h_counts, 8 * sizeof(int), cudaHostAllocDefault));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->solver.ev, cudaEventDisableTiming));
    }
    carve(pl, ctx->solver.blob, wave);
    // work-item tables
    {
        std::vector<int> tab;
        tab.insert(tab.end(), pl.jac_col.begin(), pl.jac_col.end());
        tab.insert(tab.end(), pl.jac_seg.begin(), pl.jac_seg.end());
        tab.insert(tab.end(), pl.col_item0.begin(), pl.col_item0.end());
        tab.insert(tab.end(), pl.col_nseg.begin(), pl.col_nseg.end());
        CUDA_TRY(ctx, cudaMemcpyAsync((void *)D.jac_col, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));      // tab is a stack-lifetime vector
    }
    D.xtol = xtol; D.epsfcn = epsfcn; D.factor = 1.0; D.maxfev = maxfev; D.run_mode = run_mode;
    D.counters = ctx->d_counters;
    D.sm_count = ctx->sm_count;
    D.phase_clocks = getenv("SOCP_PHASE_CLOCKS") ? 1 : 0;
    {
        // SOCP_JAC=lanes4 (builds with -DSOCP_JAC_LANES4 only): the one-barrier, four-lanes-per-column Householder
        // routines (qrfac_w / qform_w; same bits, measured SLOWER than the thread-per-column ones: 1074 vs 785 ms per step)
        const char *jm = getenv("SOCP_JAC");
        D.jac_fast = (jm && !strcmp(jm, "lanes4")) ? 1 : 0;
        // SOCP_JAC=old: the barrier-per-reflector routines everywhere (A/B runs); default: register-window routines
        // wherever the Jacobian has the block-bidiagonal + border structure
        D.jac_window = (jm && (!strcmp(jm, "old") || !strcmp(jm, "lanes4"))) ? 0 : 1;
    }
    const int grid_int = ctx->sm_count * 8;
    const int grid_adv = ctx->sm_count * 6;
    // debugging aid: SOCP_ROUND_LOG=<file> (with profiling on) logs every round: index, residual and
    // Jacobian requests entering it, and the CUDA-event time of its four kernels (ms)
    const char *round_log_path = ctx->profile ? getenv("SOCP_ROUND_LOG") : nullptr;
    FILE *round_log = round_log_path ? fopen(round_log_path, "a") : nullptr;
    const int check_every = (run_mode == RUN_SOLVE && !round_log) ? 8 : 1;   // <= kProfSlots
    const long max_rounds = (run_mode == RUN_SOLVE) ? (long)maxfev + 8 : (run_mode == RUN_FDJAC ? 2 : 1);

    for (long first = 0; first < B; first += wave) {
        const long Bw = std::min(wave, B - first);
        D.B = Bw;
        D.mparams = d_mparams + first * D.np;
        D.time = d_time + first * (D.M + 1);
        D.Xb = d_Xb + first * (D.M + 1) * D.dim;
        const long nthreads = Bw * D.P;
        // the wave's descriptor in global memory, for the device functions that take it by reference (pageable source:
        // the copy is staged before the call returns)
        CUDA_TRY(ctx, cudaMemcpyAsync((void *)D.self, &D, sizeof(SolverDev), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(D.counts, 0, 8 * sizeof(int), ctx->stream));
        solver_init<<<(unsigned)((nthreads + 255) / 256), 256, 0, ctx->stream>>>(D, d_x_in, first, d_active);
        ctx->launches += 1;
        int cur = 0, pending = 0;
        int entering[2] = {(int)Bw, 0};
        // live problems as of the last look at the counters: they only retire, so this bounds the work of
        // every later round; the grids follow it (every kernel strides over its work list, so any grid is
        // correct) and the tail of a solve -- a handful of live problems -- stops launching 1184 idle CTAs
        long live = (live_hint >= 0) ? std::min(Bw, std::max<long>(live_hint, 1)) : Bw;
        for (long round = 0; round < max_rounds; ++round) {
            const long items = live * std::max(D.nJ, D.P);                       // upper bound on work items
            const int gi = (int)std::max<long>(1, std::min<long>(grid_int, (items + 127) / 128));
            const int ga = (int)std::max<long>(1, std::min<long>(grid_adv, live));
            // `items` bounds the work items of this round (live problems only retire): the integrator may switch to
            // cooperative groups when they cannot fill the GPU
            launch_round_any(ctx, D, cur, gi, ga, ctx->profile ? pending : -1, items);
            if (ctx->launch_error) {                   // a launch could not be configured: nothing of this round ran
                const int code = ctx->launch_error;
                ctx->launch_error = 0;
                cudaStreamSynchronize(ctx->stream);
                if (round_log) fclose(round_log);
                return code;
            }
            ++pending;
            cur = 1 - cur;
            if (run_mode == RUN_SOLVE && (round % check_every) == check_every - 1) {
                CUDA_TRY(ctx, cudaMemcpyAsync(ctx->solver.h_counts, D.counts, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                CUDA_TRY(ctx, cudaEventRecord(ctx->solver.ev, ctx->stream));
                CUDA_TRY(ctx, cudaEventSynchronize(ctx->solver.ev));
                if (ctx->profile) prof_harvest(ctx, pending, round_log, round, entering);
                pending = 0;
                entering[0] = ctx->solver.h_counts[cur * 2]; entering[1] = ctx->solver.h_counts[cur * 2 + 1];
                live = (long)entering[0] + entering[1];
                if (ctx->solver.h_counts[cur * 2] + ctx->solver.h_counts[cur * 2 + 1] == 0) break;
            }
        }
        if (ctx->profile && pending > 0) {
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            prof_harvest(ctx, pending);
        }
        const long nfin = std::max<long>(nthreads, d_fjac_out ? Bw * (long)D.P * D.P : 0);
        solver_finish<<<(unsigned)((nfin + 255) / 256), 256, 0, ctx->stream>>>(D, first, d_x_out, d_fvec_out, d_fjac_out, d_info, d_nfev, d_fnorm, d_njev, d_active);
        ctx->launches += 1;
        CUDA_TRY(ctx, cudaGetLastError());
    }
    if (round_log) fclose(round_log);
    return SOCP_OK;
}

int check_problem_args(socp_ctx *ctx, const socp_shape *shape, long B, const void *a, const void *b, const void *c,
                       const void *d, const void *e) {
    if (!ctx) return SOCP_ERR_ARG;
    if (!shape || B < 0 || !a || !b || !c || !d || !e) return fail(ctx, SOCP_ERR_ARG, "NULL argument or negative batch size");
    if (socp_num_param(shape) < 0) return fail(ctx, SOCP_ERR_ARG, "bad shape");
    return SOCP_OK;
}

// error return after host buffers may have been staged: no asynchronous copy from (or to) a caller buffer
// is left in flight when the call returns an error
int bail(socp_ctx *ctx, int rc) {
    if (ctx) cudaStreamSynchronize(ctx->stream);
    return rc;
}

}  // namespace

extern "C" {

int socp_residual_batch(socp_ctx *ctx, const socp_shape *shape, long B, const double *mparams,
                        const double *time, const double *Xb, const double *x, double *fvec, int mem) {
    int rc = check_problem_args(ctx, shape, B, mparams, time, Xb, x, fvec);
    if (rc != SOCP_OK || B == 0) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int P = socp_num_param(shape), dim = kDim[shape->model_id], np = kNP[shape->model_id], M = shape->num_multi;
    const double *d_mp = stage_in(ctx, SLOT_MPARAMS, mparams, (size_t)B * np, mem, &rc);
    const double *d_time = stage_in(ctx, SLOT_TIME, time, (size_t)B * (M + 1), mem, &rc);
    const double *d_Xb = stage_in(ctx, SLOT_XB, Xb, (size_t)B * (M + 1) * dim, mem, &rc);
    const double *d_x = stage_in(ctx, SLOT_X, x, (size_t)B * P, mem, &rc);
    double *d_f = stage_out(ctx, SLOT_FVEC, fvec, (size_t)B * P, mem, &rc);
    if (rc != SOCP_OK) return bail(ctx, rc);
    rc = run_solver(ctx, shape, B, d_mp, d_time, d_Xb, d_x, RUN_RESIDUAL, 0., 1, 1e-15, nullptr, d_f, nullptr, nullptr, nullptr, nullptr);
    if (rc != SOCP_OK) return bail(ctx, rc);
    if ((rc = fetch_out(ctx, fvec, d_f, (size_t)B * P, mem)) != SOCP_OK) return rc;
    if (mem == SOCP_HOST) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SOCP_OK;
}

int socp_fdjac_batch(socp_ctx *ctx, const socp_shape *shape, long B, const double *mparams,
                     const double *time, const double *Xb, const double *x, double epsfcn,
                     double *fjac, int mem) {
    int rc = check_problem_args(ctx, shape, B, mparams, time, Xb, x, fjac);
    if (rc != SOCP_OK || B == 0) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int P = socp_num_param(shape), dim = kDim[shape->model_id], np = kNP[shape->model_id], M = shape->num_multi;
    const double *d_mp = stage_in(ctx, SLOT_MPARAMS, mparams, (size_t)B * np, mem, &rc);
    const double *d_time = stage_in(ctx, SLOT_TIME, time, (size_t)B * (M + 1), mem, &rc);
    const double *d_Xb = stage_in(ctx, SLOT_XB, Xb, (size_t)B * (M + 1) * dim, mem, &rc);
    const double *d_x = stage_in(ctx, SLOT_X, x, (size_t)B * P, mem, &rc);
    double *d_j = stage_out(ctx, SLOT_FJAC, fjac, (size_t)B * P * P, mem, &rc);
    if (rc != SOCP_OK) return bail(ctx, rc);
    rc = run_solver(ctx, shape, B, d_mp, d_time, d_Xb, d_x, RUN_FDJAC, 0., 1, epsfcn, nullptr, nullptr, d_j, nullptr, nullptr, nullptr);
    if (rc != SOCP_OK) return bail(ctx, rc);
    if ((rc = fetch_out(ctx, fjac, d_j, (size_t)B * P * P, mem)) != SOCP_OK) return rc;
    if (mem == SOCP_HOST) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SOCP_OK;
}

int socp_solve_batch(socp_ctx *ctx, const socp_shape *shape, long B, const double *mparams,
                     const double *time, const double *Xb, double *x, double xtol, int maxfev,
                     int *info, int *nfev, double *fnorm, int mem) {
    int rc = check_problem_args(ctx, shape, B, mparams, time, Xb, x, info);
    if (rc != SOCP_OK) return bail(ctx, rc);
    if (xtol < 0. || maxfev <= 0) {
        // hybrd's own argument check: info = 0 and x untouched (MINPACK hybrd, "check the input parameters")
        if (mem == SOCP_HOST) for (long b = 0; b < B; ++b) { info[b] = 0; if (nfev) nfev[b] = 0; }
        else if (B > 0) {      // device buffers: the same values, stream-ordered
            cudaSetDevice(ctx->device);
            cudaMemsetAsync(info, 0, sizeof(int) * (size_t)B, ctx->stream);
            if (nfev) cudaMemsetAsync(nfev, 0, sizeof(int) * (size_t)B, ctx->stream);
        }
        return fail(ctx, SOCP_ERR_ARG, "xtol < 0 or maxfev <= 0");
    }
    if (B == 0) return SOCP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int P = socp_num_param(shape), dim = kDim[shape->model_id], np = kNP[shape->model_id], M = shape->num_multi;
    const double *d_mp = stage_in(ctx, SLOT_MPARAMS, mparams, (size_t)B * np, mem, &rc);
    const double *d_time = stage_in(ctx, SLOT_TIME, time, (size_t)B * (M + 1), mem, &rc);
    const double *d_Xb = stage_in(ctx, SLOT_XB, Xb, (size_t)B * (M + 1) * dim, mem, &rc);
    const double *d_xin = stage_in(ctx, SLOT_X, (const double *)x, (size_t)B * P, mem, &rc);
    double *d_x = (mem == SOCP_DEVICE) ? x : (double *)d_xin;
    int *d_info = stage_out(ctx, SLOT_INFO, info, (size_t)B, mem, &rc);
    int *d_nfev = stage_out(ctx, SLOT_NFEV, nfev, (size_t)B, mem, &rc);
    double *d_fn = stage_out(ctx, SLOT_FNORM, fnorm, (size_t)B, mem, &rc);
    if (rc != SOCP_OK) return bail(ctx, rc);
    rc = run_solver(ctx, shape, B, d_mp, d_time, d_Xb, d_xin, RUN_SOLVE, xtol, maxfev, 1e-15, d_x, nullptr, nullptr, d_info, d_nfev, d_fn);
    if (rc != SOCP_OK) return bail(ctx, rc);
    if ((rc = fetch_out(ctx, x, d_x, (size_t)B * P, mem)) != SOCP_OK) return rc;
    if ((rc = fetch_out(ctx, info, d_info, (size_t)B, mem)) != SOCP_OK) return rc;
    if ((rc = fetch_out(ctx, nfev, d_nfev, (size_t)B, mem)) != SOCP_OK) return rc;
    if ((rc = fetch_out(ctx, fnorm, d_fn, (size_t)B, mem)) != SOCP_OK) return rc;
    if (mem == SOCP_HOST) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SOCP_OK;
}

int socp_solve_hybrj_batch(socp_ctx *ctx, const socp_shape *shape, long B, const double *mparams,
                           const double *time, const double *Xb, double *x, double xtol, int maxfev,
                           int *info, int *nfev, int *njev, double *fnorm, int mem) {
    int rc = check_problem_args(ctx, shape, B, mparams, time, Xb, x, info);
    if (rc != SOCP_OK) return bail(ctx, rc);
    if (xtol < 0. || maxfev <= 0) {
        if (mem == SOCP_HOST) for (long b = 0; b < B; ++b) { info[b] = 0; if (nfev) nfev[b] = 0; if (njev) njev[b] = 0; }
        else if (B > 0) {
            cudaSetDevice(ctx->device);
            cudaMemsetAsync(info, 0, sizeof(int) * (size_t)B, ctx->stream);
            if (nfev) cudaMemsetAsync(nfev, 0, sizeof(int) * (size_t)B, ctx->stream);
            if (njev) cudaMemsetAsync(njev, 0, sizeof(int) * (size_t)B, ctx->stream);
        }
        return fail(ctx, SOCP_ERR_ARG, "xtol < 0 or maxfev <= 0");
    }
    if (B == 0) return SOCP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int P = socp_num_param(shape), dim = kDim[shape->model_id], np = kNP[shape->model_id], M = shape->num_multi;
    const double *d_mp = stage_in(ctx, SLOT_MPARAMS, mparams, (size_t)B * np, mem, &rc);
    const double *d_time = stage_in(ctx, SLOT_TIME, time, (size_t)B * (M + 1), mem, &rc);
    const double *d_Xb = stage_in(ctx, SLOT_XB, Xb, (size_t)B * (M + 1) * dim, mem, &rc);
    const double *d_xin = stage_in(ctx, SLOT_X, (const double *)x, (size_t)B * P, mem, &rc);
    double *d_x = (mem == SOCP_DEVICE) ? x : (double *)d_xin;
    int *d_info = stage_out(ctx, SLOT_INFO, info, (size_t)B, mem, &rc);
    int *d_nfev = stage_out(ctx, SLOT_NFEV, nfev, (size_t)B, mem, &rc);
    int *d_njev = stage_out(ctx, SLOT_AUX0, njev, (size_t)B, mem, &rc);
    double *d_fn = stage_out(ctx, SLOT_FNORM, fnorm, (size_t)B, mem, &rc);
    if (rc != SOCP_OK) return bail(ctx, rc);
    rc = run_solver(ctx, shape, B, d_mp, d_time, d_Xb, d_xin, RUN_SOLVE, xtol, maxfev, 1e-15, d_x, nullptr, nullptr, d_info, d_nfev, d_fn, 1, d_njev);
    if (rc != SOCP_OK) return bail(ctx, rc);
    if ((rc = fetch_out(ctx, x, d_x, (size_t)B * P, mem)) != SOCP_OK) return rc;
    if ((rc = fetch_out(ctx, info, d_info, (size_t)B, mem)) != SOCP_OK) return rc;
    if ((rc = fetch_out(ctx, nfev, d_nfev, (size_t)B, mem)) != SOCP_OK) return rc;
    if ((rc = fetch_out(ctx, njev, d_njev, (size_t)B, mem)) != SOCP_OK) return rc;
    if ((rc = fetch_out(ctx, fnorm, d_fn, (size_t)B, mem)) != SOCP_OK) return rc;
    if (mem == SOCP_HOST) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SOCP_OK;
}

int socp_jacobian_batch(socp_ctx *ctx, const socp_shape *shape, long B, const double *mparams,
                        const double *time, const double *Xb, const double *x, double *fjac, int mem) {
    int rc = check_problem_args(ctx, shape, B, mparams, time, Xb, x, fjac);
    if (rc != SOCP_OK || B == 0) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int P = socp_num_param(shape), dim = kDim[shape->model_id], np = kNP[shape->model_id], M = shape->num_multi;
    const double *d_mp = stage_in(ctx, SLOT_MPARAMS, mparams, (size_t)B * np, mem, &rc);
    const double *d_time = stage_in(ctx, SLOT_TIME, time, (size_t)B * (M + 1), mem, &rc);
    const double *d_Xb = stage_in(ctx, SLOT_XB, Xb, (size_t)B * (M + 1) * dim, mem, &rc);
    const double *d_x = stage_in(ctx, SLOT_X, x, (size_t)B * P, mem, &rc);
    double *d_j = stage_out(ctx, SLOT_FJAC, fjac, (size_t)B * P * P, mem, &rc);
    if (rc != SOCP_OK) return bail(ctx, rc);
    rc = run_solver(ctx, shape, B, d_mp, d_time, d_Xb, d_x, RUN_FDJAC, 0., 1, 1e-15, nullptr, nullptr, d_j, nullptr, nullptr, nullptr, 1, nullptr);
    if (rc != SOCP_OK) return bail(ctx, rc);
    if ((rc = fetch_out(ctx, fjac, d_j, (size_t)B * P * P, mem)) != SOCP_OK) return rc;
    if (mem == SOCP_HOST) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SOCP_OK;
}

int socp_traj_var_batch(socp_ctx *ctx, int model_id, int step_nbr, long B, const double *mparams,
                        const double *t0, const double *tf, const double *X0, double *Xf, int mem) {
    if (!ctx) return SOCP_ERR_ARG;
    if (model_id < 0 || model_id >= SOCP_NUM_MODELS || B < 0 || !mparams || !t0 || !tf || !X0 || !Xf)
        return fail(ctx, SOCP_ERR_ARG, "socp_traj_var_batch: bad arguments");
    if (model_id != SOCP_DOUBLE_INTEGRATOR)
        return fail(ctx, SOCP_ERR_UNSUPPORTED, "variational integration: only the double integrator implements it, as in the reference");
    if (B == 0) return SOCP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int N = 2 * kDim[model_id], NV = N * (N + 1), np = kNP[model_id];
    const int S = step_nbr > 0 ? step_nbr : kSteps[model_id];
    int rc = SOCP_OK;
    const double *d_mp = stage_in(ctx, SLOT_MPARAMS, mparams, (size_t)B * np, mem, &rc);
    const double *d_t0 = stage_in(ctx, SLOT_T0, t0, (size_t)B, mem, &rc);
    const double *d_tf = stage_in(ctx, SLOT_TF, tf, (size_t)B, mem, &rc);
    const double *d_X0 = stage_in(ctx, SLOT_X0, X0, (size_t)B * NV, mem, &rc);
    double *d_Xf = stage_out(ctx, SLOT_XF, Xf, (size_t)B * NV, mem, &rc);
    if (rc != SOCP_OK) return bail(ctx, rc);
    const long threads = B * 16;
    traj_var_kernel<DOUBLE_INTEGRATOR><<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(B, S, d_mp, d_t0, d_tf, d_X0, d_Xf, ctx->d_counters);
    ctx->launches += 1;
    CUDA_TRY(ctx, cudaGetLastError());
    if ((rc = fetch_out(ctx, Xf, d_Xf, (size_t)B * NV, mem)) != SOCP_OK) return rc;
    if (mem == SOCP_HOST) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SOCP_OK;
}

// ---- continuation: per-problem homotopy state machines ON THE DEVICE --------------------------------------------
// Each problem runs the reference's loop (shooting.cpp:598-692 / :695-778) with its own b, b_prec; all problems
// still in their loop are solved together, one batched solve per pass.  The state machine, the boundary / parameter
// data of the next pass and the accepted solutions never leave the device: per pass the host launches four kernels
// around run_solver and reads one int (problems still active).  SOCP_HOST buffers are staged once before the first
// pass and fetched once after the last.
static int run_continuation(socp_ctx *ctx, const socp_shape *shape, long B, double *mparams, const double *time_prec,
                            const double *Xb_prec, const double *time_des, const double *Xb_des, double *x, double xtol,
                            int maxfev, double step, double step_min, int param_idx, const double *goal, int *info,
                            int *calls, int mem) {
    const int P = socp_num_param(shape), dim = kDim[shape->model_id], np = kNP[shape->model_id], M = shape->num_multi;
    const int nt = M + 1, nx = (M + 1) * dim;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = SOCP_OK;
    // caller arrays on the device (staged copies in HOST mode)
    double *d_mp = (double *)stage_in(ctx, SLOT_MPARAMS, (const double *)mparams, (size_t)B * np, mem, &rc);
    const double *d_tp = stage_in(ctx, SLOT_TIME, time_prec, (size_t)B * nt, mem, &rc);
    const double *d_Xp = stage_in(ctx, SLOT_XB, Xb_prec, (size_t)B * nx, mem, &rc);
    const double *d_td = stage_in(ctx, SLOT_CONT_A, time_des, (size_t)B * nt, mem, &rc);
    const double *d_Xd = stage_in(ctx, SLOT_CONT_B, Xb_des, (size_t)B * nx, mem, &rc);
    const double *d_goal = stage_in(ctx, SLOT_AUX0, goal, (size_t)B, mem, &rc);
    double *d_x = (double *)stage_in(ctx, SLOT_X, (const double *)x, (size_t)B * P, mem, &rc);
    int *d_info = stage_out(ctx, SLOT_INFO, info, (size_t)B, mem, &rc);
    int *d_calls = stage_out(ctx, SLOT_NFEV, calls, (size_t)B * 2, mem, &rc);
    if (rc != SOCP_OK) return bail(ctx, rc);
    // homotopy state
    const bool boundary = param_idx < 0;
    size_t need = 0;
    {
        Carver c(nullptr);
        c.take<double>(3 * (size_t)B); c.take<double>((size_t)B * P);
        if (boundary) { c.take<double>((size_t)B * nt); c.take<double>((size_t)B * nx); }
        c.take<int>(3 * (size_t)B + 8);
        need = c.off + 512;
    }
    void *blob = ws(ctx, SLOT_CONT, need);
    if (!blob) return bail(ctx, SOCP_ERR_NOMEM);
    Carver c(blob);
    ContDev C;
    memset(&C, 0, sizeof C);
    C.B = B; C.P = P; C.np = np; C.nt = nt; C.nx = nx; C.param_idx = param_idx; C.step = step; C.step_min = step_min;
    double *st3 = c.take<double>(3 * (size_t)B);
    C.b = st3; C.b_prec = st3 + B; C.rstart = st3 + 2 * B;
    C.xw = c.take<double>((size_t)B * P);
    if (boundary) { C.time_w = c.take<double>((size_t)B * nt); C.Xb_w = c.take<double>((size_t)B * nx); }
    int *ints = c.take<int>(3 * (size_t)B + 8);
    C.active = ints; C.info_w = ints + B; C.nfev_w = ints + 2 * B; C.n_active = ints + 3 * B;
    C.goal = d_goal; C.x = d_x; C.mparams = d_mp;
    C.time_prec = d_tp; C.Xb_prec = d_Xp; C.time_des = d_td; C.Xb_des = d_Xd;
    C.info = d_info; C.calls = d_calls;
    if (!ctx->solver.h_counts) {
        CUDA_TRY(ctx, cudaHostAlloc((void **)&ctx->solver.h_counts, 8 * sizeof(int), cudaHostAllocDefault));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->solver.ev, cudaEventDisableTiming));
    }
    const unsigned gP = (unsigned)(((size_t)B * P + 255) / 256);
    cont_begin<<<gP, 256, 0, ctx->stream>>>(C);
    ctx->launches += 1;
    long n_active = B;
    // the loop ends when no problem is left in its homotopy; every pass retires or advances each active problem,
    // and a problem needs at most ~ 1/step + 2 log2(1/step_min) passes
    for (long pass = 0; n_active > 0; ++pass) {
        const size_t items = boundary ? (size_t)B * (nt + nx) : (size_t)B;
        cont_setup<<<(unsigned)((items + 255) / 256), 256, 0, ctx->stream>>>(C);
        ctx->launches += 1;
        rc = run_solver(ctx, shape, B, d_mp, boundary ? C.time_w : d_tp, boundary ? C.Xb_w : d_Xp, C.xw, RUN_SOLVE, xtol, maxfev,
                        1e-15, C.xw, nullptr, nullptr, C.info_w, C.nfev_w, nullptr, 0, nullptr, C.active, n_active);
        if (rc != SOCP_OK) return bail(ctx, rc);
        cont_update<<<(unsigned)(((size_t)B * 32 + 255) / 256), 256, 0, ctx->stream>>>(C);
        ctx->launches += 1;
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->solver.h_counts + 4, C.n_active, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        n_active = ctx->solver.h_counts[4];
    }
    CUDA_TRY(ctx, cudaGetLastError());
    if ((rc = fetch_out(ctx, x, (const double *)d_x, (size_t)B * P, mem)) != SOCP_OK) return bail(ctx, rc);
    if (!boundary && (rc = fetch_out(ctx, mparams, (const double *)d_mp, (size_t)B * np, mem)) != SOCP_OK) return bail(ctx, rc);
    if ((rc = fetch_out(ctx, info, (const int *)d_info, (size_t)B, mem)) != SOCP_OK) return bail(ctx, rc);
    if ((rc = fetch_out(ctx, calls, (const int *)d_calls, (size_t)B * 2, mem)) != SOCP_OK) return bail(ctx, rc);
    if (mem == SOCP_HOST) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SOCP_OK;
}

int socp_continuation_param_batch(socp_ctx *ctx, const socp_shape *shape, long B, double *mparams,
                                  const double *time, const double *Xb, double *x, double xtol,
                                  int maxfev, double step, int param_idx, const double *goal,
                                  double step_min, int *info, int *calls, int mem) {
    int rc = check_problem_args(ctx, shape, B, mparams, time, Xb, x, info);
    if (rc != SOCP_OK) return rc;
    const int np = kNP[shape->model_id];
    if (param_idx < 0 || param_idx >= np || !goal) return fail(ctx, SOCP_ERR_ARG, "bad parameter index / goal");
    if (xtol < 0. || maxfev <= 0) return fail(ctx, SOCP_ERR_ARG, "xtol < 0 or maxfev <= 0");
    if (step <= 0) step = 1.0;                              // shooting.cpp:352-355
    if (B == 0) return SOCP_OK;
    return run_continuation(ctx, shape, B, mparams, time, Xb, nullptr, nullptr, x, xtol, maxfev, step, step_min, param_idx,
                            goal, info, calls, mem);
}

int socp_continuation_boundary_batch(socp_ctx *ctx, const socp_shape *shape, long B,
                                     const double *mparams, const double *time_prec,
                                     const double *Xb_prec, const double *time_des,
                                     const double *Xb_des, double *x, double xtol, int maxfev,
                                     double step, double step_min, int *info, int *calls, int mem) {
    int rc = check_problem_args(ctx, shape, B, mparams, time_prec, Xb_prec, x, info);
    if (rc != SOCP_OK) return rc;
    if (!time_des || !Xb_des) return fail(ctx, SOCP_ERR_ARG, "NULL desired boundary data");
    if (xtol < 0. || maxfev <= 0) return fail(ctx, SOCP_ERR_ARG, "xtol < 0 or maxfev <= 0");
    if (step <= 0) return fail(ctx, SOCP_ERR_ARG, "continuation step must be > 0 (use socp_solve_batch otherwise)");
    if (B == 0) return SOCP_OK;
    return run_continuation(ctx, shape, B, (double *)mparams, time_prec, Xb_prec, time_des, Xb_des, x, xtol, maxfev, step,
                            step_min, -1, nullptr, info, calls, mem);
}

}  // extern "C"
