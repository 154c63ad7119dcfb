// oracle/ref_bridge.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Plain-C handle API over the UNMODIFIED reference classes (compiled from /root/reference by
// oracle/Makefile into oracle/_ref/libsocp_ref.so) so that Python tests / bench.py can drive the
// real `model` / `shooting` code: trajectories (model::ComputeTraj, model.hpp:77), the shooting
// residual (shooting::StaticShootingFunction, shooting.cpp:859, reached through the normal
// SolveOCP -> hybrd call path: this file owns the `hybrd`/`hybrj` symbols and can either forward
// to the clean-room solver (oracle/minpack.c) or just evaluate the callback), and full solves /
// continuations (shooting::SolveOCP, shooting.cpp:315-362).
//
// Nothing here is reference code; the reference headers are only #included.
#include <iostream>
#include <vector>
#include <string>
#include <cstring>
#include <cmath>

#include "minpack.h"
#include "socp/shooting.hpp"
#include "models/goddard/goddard.hpp"
#include "models/doubleIntegrator/doubleIntegrator.hpp"
#include "models/covid19/covid19.hpp"
#include "models/vtolUAV/vtolUAV.hpp"
#include "models/interceptor/interceptor.hpp"
#include "maps/obstacle/obstacle.hpp"

extern "C" {
// the clean-room solver, compiled from oracle/minpack.c with -Dhybrd=mp_core_hybrd etc.
int mp_core_hybrd(minpack_func_nn, void*, int, double*, double*, double, int, int, int, double,
                  double*, int, double, int, int*, double*, int, double*, int, double*, double*,
                  double*, double*, double*);
int mp_core_hybrj(minpack_funcder_nn, void*, int, double*, double*, double*, int, double, int,
                  double*, int, double, int, int*, int*, double*, int, double*, double*, double*,
                  double*, double*);
}

namespace {
// interception state (single-threaded use from ctypes)
enum { MODE_SOLVE = 0, MODE_RESIDUAL = 1, MODE_JACOBIAN = 2 };
int g_mode = MODE_SOLVE;
const double* g_eval_x = 0;   // where to evaluate in MODE_RESIDUAL / MODE_JACOBIAN
double* g_eval_out = 0;
struct CallLog { int info, nfev, njev, n; };
std::vector<CallLog> g_log;

enum { M_GODDARD = 0, M_DI = 1, M_COVID = 2, M_VTOL = 3, M_INTERCEPTOR = 4 };

struct ModelBox {
    int id;
    model* m;
    obstacle* obs;
    ModelBox() : id(-1), m(0), obs(0) {}
};

double* param_ref(ModelBox* b, int idx)
{
    switch (b->id) {
    case M_GODDARD: {
        static const char* names[] = {"C", "b", "KD", "kr", "u_max", "mu1", "mu2", "singularControl"};
        if (idx < 0 || idx >= 8) return 0;
        return &static_cast<goddard*>(b->m)->GetParameterDataName(names[idx]);
    }
    case M_DI: {
        doubleIntegrator::parameters_struct& p = static_cast<doubleIntegrator*>(b->m)->GetParameterData();
        double* t[] = {&p.u_max, &p.a_max, &p.muT};
        return (idx >= 0 && idx < 3) ? t[idx] : 0;
    }
    case M_COVID: {
        covid19::parameters_struct& p = static_cast<covid19*>(b->m)->GetParameterData();
        double* t[] = {&p.R0, &p.Tinf, &p.Tinc, &p.N, &p.Imax, &p.muI, &p.umin, &p.umax};
        return (idx >= 0 && idx < 8) ? t[idx] : 0;
    }
    case M_VTOL: {
        vtolUAV::parameters_struct& p = static_cast<vtolUAV*>(b->m)->GetParameterData();
        obstacle::parameters_struct& o = b->obs->GetParameterData();
        double* t[] = {&p.u_max, &p.a_max, &p.alphaT, &p.alphaV, &p.invSigmaXwp, &p.Vd, &p.ca, 0, 0,
                       &o.phiObs, &o.psiWP, &o.muObs, &o.sigmaWP};
        return (idx >= 0 && idx < 13) ? t[idx] : 0;
    }
    case M_INTERCEPTOR: {
        interceptor::parameters_struct& p = static_cast<interceptor*>(b->m)->GetParameterData();
        double* t[] = {&p.c0, &p.hr, &p.d0, &p.eta, &p.propellant_mass, &p.empty_mass, &p.q, &p.ve,
                       &p.alpha_max, &p.u_max, &p.a_max, &p.r_2p, &p.t_2p, &p.mu_gft, &p.muT,
                       &p.muV, &p.muC};
        return (idx >= 0 && idx < 17) ? t[idx] : 0;
    }
    }
    return 0;
}
}  // namespace

extern "C" {

// ---- the two symbols shooting.o needs -------------------------------------------------------
int hybrd(minpack_func_nn fcn, void* p, int n, double* x, double* fvec, double xtol, int maxfev,
          int ml, int mu, double epsfcn, double* diag, int mode, double factor, int nprint,
          int* nfev, double* fjac, int ldfjac, double* r, int lr, double* qtf, double* wa1,
          double* wa2, double* wa3, double* wa4)
{
    if (g_mode == MODE_RESIDUAL) {
        fcn(p, n, g_eval_x, g_eval_out, 1);
        *nfev = 1;
        return -999;   // != 1, so SOCP leaves tab_param untouched (shooting.cpp:588)
    }
    int info = mp_core_hybrd(fcn, p, n, x, fvec, xtol, maxfev, ml, mu, epsfcn, diag, mode, factor,
                             nprint, nfev, fjac, ldfjac, r, lr, qtf, wa1, wa2, wa3, wa4);
    CallLog c = {info, *nfev, 0, n};
    g_log.push_back(c);
    return info;
}

int hybrj(minpack_funcder_nn fcn, void* p, int n, double* x, double* fvec, double* fjac, int ldfjac,
          double xtol, int maxfev, double* diag, int mode, double factor, int nprint, int* nfev,
          int* njev, double* r, int lr, double* qtf, double* wa1, double* wa2, double* wa3,
          double* wa4)
{
    if (g_mode == MODE_RESIDUAL) {
        fcn(p, n, g_eval_x, g_eval_out, 0, n, 1);
        *nfev = 1; *njev = 0;
        return -999;
    }
    if (g_mode == MODE_JACOBIAN) {   // column-major n x n into g_eval_out
        std::vector<double> f(n);
        fcn(p, n, g_eval_x, f.data(), g_eval_out, n, 2);
        *nfev = 0; *njev = 1;
        return -999;
    }
    int info = mp_core_hybrj(fcn, p, n, x, fvec, fjac, ldfjac, xtol, maxfev, diag, mode, factor,
                             nprint, nfev, njev, r, lr, qtf, wa1, wa2, wa3, wa4);
    CallLog c = {info, *nfev, *njev, n};
    g_log.push_back(c);
    return info;
}

// ---- models ---------------------------------------------------------------------------------
void* ref_model_new(int model_id, int model_order, int step_nbr, const char* obstacle_file,
                    const char* wp_file)
{
    ModelBox* b = new ModelBox;
    b->id = model_id;
    switch (model_id) {
    case M_GODDARD: b->m = new goddard(std::string(""), step_nbr > 0 ? step_nbr : 10); break;
    case M_DI: b->m = new doubleIntegrator(model_order, std::string("")); break;
    case M_COVID: b->m = new covid19(std::string("")); break;
    case M_VTOL:
        b->obs = new obstacle(std::string(obstacle_file ? obstacle_file : ""),
                              std::string(wp_file ? wp_file : ""));
        b->m = new vtolUAV(*b->obs, std::string(""));
        break;
    case M_INTERCEPTOR: b->m = new interceptor(std::string("")); break;
    default: delete b; return 0;
    }
    return b;
}
void ref_model_free(void* h)
{
    ModelBox* b = (ModelBox*)h;
    delete b->m;
    delete b->obs;
    delete b;
}
int ref_model_dim(void* h) { return ((ModelBox*)h)->m->GetDim(); }
int ref_model_set_param(void* h, int idx, double v)
{
    ModelBox* b = (ModelBox*)h;
    if (b->id == M_VTOL && (idx == 7 || idx == 8)) {
        vtolUAV::parameters_struct& p = static_cast<vtolUAV*>(b->m)->GetParameterData();
        (idx == 7 ? p.nWP_tot : p.nWP) = (int)v;
        return 0;
    }
    double* r = param_ref(b, idx);
    if (!r) return -1;
    *r = v;
    return 0;
}
double ref_model_get_param(void* h, int idx)
{
    ModelBox* b = (ModelBox*)h;
    if (b->id == M_VTOL && (idx == 7 || idx == 8)) {
        vtolUAV::parameters_struct& p = static_cast<vtolUAV*>(b->m)->GetParameterData();
        return idx == 7 ? p.nWP_tot : p.nWP;
    }
    double* r = param_ref(b, idx);
    return r ? *r : NAN;
}
void ref_model_set_ode_tol(void* h, double tol) { ((ModelBox*)h)->m->SetODEIntPrecision(tol); }
void ref_model_switching_times(void* h, int n, const double* ts)
{
    ((ModelBox*)h)->m->SwitchingTimesUpdate(std::vector<real>(ts, ts + n));
}
// model::ComputeTraj (model.hpp:77; interceptor.cpp:165)
void ref_traj(void* h, double t0, const double* X0, int nX, double tf, int isJac, double* Xf)
{
    model::mstate X(X0, X0 + nX);
    model::mstate Y = ((ModelBox*)h)->m->ComputeTraj(t0, X, tf, 0, isJac);
    for (size_t i = 0; i < Y.size(); ++i) Xf[i] = Y[i];
}
// odeTools::Model (RHS), model::Control, model::Hamiltonian at a point
int ref_rhs(void* h, double t, const double* X, int nX, int isJac, double* out)
{
    model::mstate Y = ((ModelBox*)h)->m->Model(t, model::mstate(X, X + nX), isJac);
    for (size_t i = 0; i < Y.size(); ++i) out[i] = Y[i];
    return (int)Y.size();
}
int ref_control(void* h, double t, const double* X, int nX, double* out)
{
    model::mcontrol Y = ((ModelBox*)h)->m->Control(t, model::mstate(X, X + nX));
    for (size_t i = 0; i < Y.size(); ++i) out[i] = Y[i];
    return (int)Y.size();
}
int ref_hamiltonian(void* h, double t, const double* X, int nX, int isJac, double* out)
{
    model::mstate Y = ((ModelBox*)h)->m->Hamiltonian(t, model::mstate(X, X + nX), isJac);
    for (size_t i = 0; i < Y.size(); ++i) out[i] = Y[i];
    return (int)Y.size();
}
void ref_obstacle_eval(void* h, const double* pos, double* func, double* grad)
{
    ModelBox* b = (ModelBox*)h;
    std::vector<real> p(pos, pos + 3), g(3);
    real f = 0;
    b->obs->Function(p, f);
    b->obs->Gradient(p, g);
    *func = f;
    grad[0] = g[0]; grad[1] = g[1]; grad[2] = g[2];
}
void ref_interceptor_init_analytical(void* h, double ti, double* Xi, double tf, double* Xf)
{
    model::mstate a(Xi, Xi + 12), b(Xf, Xf + 12);
    static_cast<interceptor*>(((ModelBox*)h)->m)->InitAnalytical(ti, a, tf, b);
    for (int i = 0; i < 12; ++i) { Xi[i] = a[i]; Xf[i] = b[i]; }
}

// ---- shooting -------------------------------------------------------------------------------
void* ref_shooting_new(void* model_h, int numMulti, int numThread)
{
    return new shooting(*((ModelBox*)model_h)->m, numMulti, numThread);
}
void ref_shooting_free(void* s) { delete (shooting*)s; }
void ref_shooting_resize(void* s, int numMulti, int numThread) { ((shooting*)s)->Resize(numMulti, numThread); }
void ref_shooting_set_precision(void* s, double xtol) { ((shooting*)s)->SetPrecision(xtol); }
void ref_shooting_set_cont_min_step(void* s, double v) { ((shooting*)s)->SetContinuationMinStep(v); }
void ref_shooting_set_mode_final(void* s, int mode_tf, const int* mode_Xf, int dim)
{
    ((shooting*)s)->SetMode(mode_tf, std::vector<int>(mode_Xf, mode_Xf + dim));
}
void ref_shooting_set_mode(void* s, const int* mode_t, const int* mode_X, int nodes, int dim)
{
    std::vector<int> mt(mode_t, mode_t + nodes);
    std::vector<std::vector<int> > mx(nodes);
    for (int i = 0; i < nodes; ++i) mx[i] = std::vector<int>(mode_X + i * dim, mode_X + (i + 1) * dim);
    ((shooting*)s)->SetMode(mt, mx);
}
void ref_shooting_init(void* s, double ti, const double* Xi, double tf, const double* Xf, int nX)
{
    ((shooting*)s)->InitShooting(ti, model::mstate(Xi, Xi + nX), tf, model::mstate(Xf, Xf + nX));
}
void ref_shooting_init_v(void* s, const double* vt, const double* vX, int nodes, int nX)
{
    std::vector<real> t(vt, vt + nodes);
    std::vector<model::mstate> X(nodes);
    for (int i = 0; i < nodes; ++i) X[i] = model::mstate(vX + i * nX, vX + (i + 1) * nX);
    ((shooting*)s)->InitShooting(t, X);
}
void ref_shooting_desired(void* s, double ti, const double* Xi, double tf, const double* Xf, int nX)
{
    ((shooting*)s)->SetDesiredState(ti, model::mstate(Xi, Xi + nX), tf, model::mstate(Xf, Xf + nX));
}
void ref_shooting_desired_v(void* s, const double* vt, const double* vX, int nodes, int nX)
{
    std::vector<real> t(vt, vt + nodes);
    std::vector<model::mstate> X(nodes);
    for (int i = 0; i < nodes; ++i) X[i] = model::mstate(vX + i * nX, vX + (i + 1) * nX);
    ((shooting*)s)->SetDesiredState(t, X);
}
int ref_shooting_solve(void* s, double step)
{
    g_mode = MODE_SOLVE;
    return ((shooting*)s)->SolveOCP(step);
}
int ref_shooting_solve_param(void* s, void* model_h, double step, int param_idx, double goal)
{
    g_mode = MODE_SOLVE;
    double* r = param_ref((ModelBox*)model_h, param_idx);
    if (!r) return -1000;
    return ((shooting*)s)->SolveOCP(step, *r, goal);
}
int ref_shooting_num_param(void* s)
{
    std::vector<real> v;
    ((shooting*)s)->GetParameters(v);
    return (int)v.size();
}
void ref_shooting_get_params(void* s, double* out)
{
    std::vector<real> v;
    ((shooting*)s)->GetParameters(v);
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
}
void ref_shooting_get_solution(void* s, int nodes, int nX, double* vt, double* vX)
{
    std::vector<real> t(nodes);
    std::vector<model::mstate> X(nodes, model::mstate(nX));
    ((shooting*)s)->GetSolution(t, X);
    for (int i = 0; i < nodes; ++i) {
        vt[i] = t[i];
        for (int k = 0; k < nX; ++k) vX[i * nX + k] = X[i][k];
    }
}
void ref_shooting_move(void* s, double tf, int nX, double* out)
{
    model::mstate X = ((shooting*)s)->Move(tf);
    for (int k = 0; k < nX && k < (int)X.size(); ++k) out[k] = X[k];
}
void ref_shooting_call_number(void* s, int* out)
{
    std::vector<int> v = ((shooting*)s)->GetCallNumber();
    out[0] = v[0]; out[1] = v[1];
}
// residual F(x) exactly as hybrd would see it on SolveOCP(0.0) (boundary data = desired data)
void ref_shooting_residual(void* s, const double* x, double* fvec)
{
    g_mode = MODE_RESIDUAL;
    g_eval_x = x; g_eval_out = fvec;
    ((shooting*)s)->SolveOCP(0.0);
    g_mode = MODE_SOLVE;
}
// analytic Jacobian (modelOrder==1 models), column-major P x P
void ref_shooting_jacobian(void* s, const double* x, double* fjac)
{
    g_mode = MODE_JACOBIAN;
    g_eval_x = x; g_eval_out = fjac;
    ((shooting*)s)->SolveOCP(0.0);
    g_mode = MODE_SOLVE;
}

// ---- solver-call log ------------------------------------------------------------------------
void ref_log_clear() { g_log.clear(); }
int ref_log_size() { return (int)g_log.size(); }
void ref_log_get(int i, int* out4)
{
    out4[0] = g_log[i].info; out4[1] = g_log[i].nfev; out4[2] = g_log[i].njev; out4[3] = g_log[i].n;
}

}  // extern "C"
