/*
 * oracle/socp_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C CPU restatement of the reference's shooting hot path (bherisse/socp):
 * RK4 integration of the state/costate ODE, the five shipped models, the multiple-shooting
 * residual, and the Powell-hybrid solve + continuation loops around it.  Every function cites
 * the reference file:line it follows (paths relative to /root/reference).
 *
 * PARITY PIN: checked bit-for-bit against the unmodified reference compiled into
 * oracle/_ref/libsocp_ref.so (tests/test_oracle_vs_ref.py, run in the authoring container)
 * and against the golden vectors that script committed under tests/golden/.
 */
#ifndef SOCP_ORACLE_H
#define SOCP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { SO_GODDARD = 0, SO_DI = 1, SO_COVID19 = 2, SO_VTOL = 3, SO_INTERCEPTOR = 4 };
enum { SO_FIXED = 0, SO_FREE = 1, SO_CONTINUOUS = 2 };   /* model.hpp:34-38 */

#define SO_MAX_DIM 7
#define SO_MAX_N 14          /* 2*dim */
#define SO_MAX_NODES 64
#define SO_MAX_OBS 32

/* obstacle table (src/maps/obstacle/obstacle.cpp:24-36) */
typedef struct {
    int n;
    double type[SO_MAX_OBS];
    double pos[SO_MAX_OBS][3];
    double rad[SO_MAX_OBS][3];
} so_obstacles;

/* One optimal-control problem = reference `model` + `shooting::data_struct` (shooting.cpp:21-54).
 * Model parameter blocks (mparams) are indexed as in oracle/pyref.py PARAMS. */
typedef struct {
    int model_id, dim, num_multi, step_nbr;
    int mode_t[SO_MAX_NODES];               /* per node: FIXED / FREE / CONTINUOUS */
    int mode_X[SO_MAX_NODES][SO_MAX_DIM];   /* per node and state component */
    double mparams[20];
    double time[SO_MAX_NODES];              /* data->time */
    double Xb[SO_MAX_NODES][SO_MAX_DIM];    /* data->X[i][0..dim): boundary / waypoint states */
    const so_obstacles *obs;                /* vtolUAV only */
    /* hidden model state of the reference, made explicit */
    double sw[SO_MAX_NODES];                /* goddard/vtol switching times (goddard.cpp:373) */
    int nsw;
    int chart, stage;                       /* interceptor.cpp:27-29 */
    /* statistics */
    long rk4_steps;
    /* conditioning probe (tests only): when noise_ulps > 0 every RHS / Hamiltonian evaluation
     * happens at inputs multiplied component-wise by (1 + noise_ulps * 2^-53 * u), u uniform in
     * [-1,1], and every output is multiplied likewise -- a stochastic-arithmetic (backward error)
     * estimate of how far a libm / FMA difference of that many ulps can move a result */
    double noise_ulps;
    unsigned long long noise_state;
    /* integrator of model::ModelInt: 0 = fixed-step RK4 (the reference's default build), 1 = adaptive
     * Dormand-Prince as the reference's -D_USE_BOOST build (odeTools.cpp:131-134), abs = rel = ode_tol */
    int integrator;
    double ode_tol;
    long dopri_steps, dopri_rejected;       /* statistics */
} so_problem;

void so_problem_init(so_problem *p, int model_id, int num_multi);   /* model ctor defaults */
int so_num_param(const so_problem *p);                              /* shooting.cpp:179,196 */
int so_default_steps(int model_id);

void so_rhs(so_problem *p, double t, const double *X, double *dX);
int so_control(so_problem *p, double t, const double *X, double *u);
double so_hamiltonian(so_problem *p, double t, const double *X);
void so_obstacle_eval(const so_problem *p, const double *pos, double *func, double *grad);

void so_rk4_step(so_problem *p, double t, double *X, double h);                   /* odeTools.cpp:89 */
void so_integrate(so_problem *p, double *X, double t0, double tf, double dt);     /* odeTools.cpp:128 */
void so_traj(so_problem *p, double t0, const double *X0, double tf, double *Xf);  /* model.hpp:77 */
/* Boost.Odeint integrate_adaptive(make_dense_output<runge_kutta_dopri5>(tol, tol), f, X, t0, tf, dt)
 * restated from the published Boost 1.7x sources (Boost is absent here: PARITY UNPINNED, see the
 * function's comment in socp_oracle.c) */
void so_integrate_adaptive(so_problem *p, double *X, double t0, double tf, double dt, double tol);

void so_timeline(so_problem *p, const double *x, double *tl);                     /* shooting.cpp:1579 */
void so_residual(so_problem *p, const double *x, double *fvec);                   /* shooting.cpp:918 */
void so_fdjac(so_problem *p, const double *x, double epsfcn, double *fjac);       /* column-major */
/* analytic-Jacobian path (modelOrder == 1; the double integrator only, as in the reference) */
int so_has_variational(const so_problem *p);
void so_traj_var(so_problem *p, double t0, const double *X0, double tf, double *Xf);   /* (2n+1) 2n doubles */
int so_jacobian(so_problem *p, const double *x, double *fjac);                         /* shooting.cpp:996-1130, column-major */
int so_solve_hybrj(so_problem *p, double *x, double xtol, int maxfev, int *nfev, int *njev, double *fnorm);

/* SolveShootingFunction (shooting.cpp:781): hybrd with SOCP's settings. Returns info. */
int so_solve(so_problem *p, double *x, double xtol, int maxfev, int *nfev, double *fnorm);

/* SolveShootingContinuation on a model parameter (shooting.cpp:695-778).  x = tab_param in/out.
 * calls[0] = number of solver calls, calls[1] = total nfev.  Returns last info. */
int so_continuation_param(so_problem *p, double *x, double xtol, int maxfev, double step,
                          int param_idx, double goal, double step_min, int *calls);

/* SolveShootingContinuation on boundary data (shooting.cpp:598-692): homotopy from
 * (time_prec, X_prec) to (timed, Xd); arrays are [(M+1)] and [(M+1)][SO_MAX_DIM]. */
int so_continuation_boundary(so_problem *p, double *x, double xtol, int maxfev, double step,
                             const double *time_prec, const double (*X_prec)[SO_MAX_DIM],
                             const double *timed, const double (*Xd)[SO_MAX_DIM],
                             double step_min, int *calls);

#ifdef __cplusplus
}
#endif
#endif
