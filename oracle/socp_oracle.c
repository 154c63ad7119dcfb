/*
 * oracle/socp_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the reference's shooting hot path; see socp_oracle.h.
 * Expression order follows the reference so that, built without FMA contraction, results are
 * bit-identical to oracle/_ref (the unmodified reference) -- tests/test_oracle_vs_ref.py.
 * Paths in comments are relative to /root/reference.
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include "socp_oracle.h"
#include "minpack.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---- parameter block indices (same order as oracle/pyref.py PARAMS) ---------------------- */
enum { G_C, G_b, G_KD, G_kr, G_umax, G_mu1, G_mu2, G_sing };
enum { D_umax, D_amax, D_muT };
enum { C_R0, C_Tinf, C_Tinc, C_N, C_Imax, C_muI, C_umin, C_umax };
enum { V_umax, V_amax, V_alphaT, V_alphaV, V_invSigma, V_Vd, V_ca, V_nWPtot, V_nWP,
       V_phiObs, V_psiWP, V_muObs, V_sigmaWP };
enum { I_c0, I_hr, I_d0, I_eta, I_mprop, I_mempty, I_q, I_ve, I_alphamax, I_umax, I_amax,
       I_r2p, I_t2p, I_mugft, I_muT, I_muV, I_muC };
#define I_R_EARTH 6378145.0      /* interceptor.cpp:51 */
#define I_MU0 3.986e14           /* interceptor.cpp:52 */
#define I_CHART_LIMIT 0.1        /* interceptor.cpp:57 */

int so_default_steps(int model_id)
{
    /* goddard.cpp:23 (ctor arg, tests use 10), doubleIntegrator.cpp:26, covid19.cpp:38,
     * vtolUAV.cpp:38, interceptor.cpp:54 */
    static const int s[] = {10, 30, 1000, 100, 50};
    return s[model_id];
}

void so_problem_init(so_problem *p, int model_id, int num_multi)
{
    static const int dims[] = {7, 6, 4, 6, 6};
    memset(p, 0, sizeof *p);
    p->model_id = model_id;
    p->dim = dims[model_id];
    p->num_multi = num_multi;
    p->step_nbr = so_default_steps(model_id);
    p->chart = 1;
    double *m = p->mparams;
    switch (model_id) {
    case SO_GODDARD:      /* goddard.hpp:29-36 */
        m[G_C] = 3.5; m[G_b] = 7.0; m[G_KD] = 310.0; m[G_kr] = 500.0; m[G_umax] = 1.0;
        m[G_mu1] = 1.0; m[G_mu2] = 0.0; m[G_sing] = -1;
        p->sw[0] = 0.0227; p->sw[1] = 0.08; p->nsw = 2;   /* goddard.cpp:27-29 */
        break;
    case SO_DI:           /* doubleIntegrator.cpp:30-32 */
        m[D_umax] = 1; m[D_amax] = 1; m[D_muT] = 0.01;
        break;
    case SO_COVID19:      /* covid19.cpp:29-36 */
        m[C_R0] = 4; m[C_Tinf] = 10; m[C_Tinc] = 5; m[C_N] = 1; m[C_Imax] = 0.1; m[C_muI] = 1;
        m[C_umin] = -10; m[C_umax] = 20;
        break;
    case SO_VTOL:         /* vtolUAV.cpp:27-35, obstacle.cpp:45-48 */
        m[V_umax] = 10; m[V_amax] = 0.3; m[V_alphaT] = 0.05; m[V_alphaV] = 0 * 0.05;
        m[V_invSigma] = 1. / 60; m[V_Vd] = 1; m[V_ca] = 0 * 0.05; m[V_nWPtot] = 0; m[V_nWP] = 0;
        m[V_phiObs] = 1; m[V_psiWP] = 0.03; m[V_muObs] = 1; m[V_sigmaWP] = 2.5;
        break;
    case SO_INTERCEPTOR:  /* interceptor.cpp:36-50 */
        m[I_c0] = 0.00075; m[I_hr] = 7500; m[I_d0] = 0.00005; m[I_eta] = 0.442; m[I_mprop] = 200;
        m[I_mempty] = 200; m[I_q] = 10; m[I_ve] = 1500; m[I_alphamax] = M_PI / 6; m[I_umax] = 1;
        m[I_amax] = 1500; m[I_mugft] = 1; m[I_muT] = 0; m[I_muV] = 1; m[I_muC] = 0;
        break;
    }
}

int so_num_param(const so_problem *p)
{
    /* shooting.cpp:192-196 */
    int nfree = 0;
    for (int j = 0; j <= p->num_multi; ++j)
        if (p->mode_t[j] == SO_FREE) ++nfree;
    return 2 * p->dim * p->num_multi + nfree;
}

/* =========================== goddard (src/models/goddard/goddard.cpp) ==================== */

/* goddard.cpp:188-253 */
static double goddard_singular(const so_problem *p, const double *X)
{
    const double *m = p->mparams;
    double x = X[0], y = X[1], z = X[2], vx = X[3], vy = X[4], vz = X[5], mass = X[6];
    double p_x = X[7], p_y = X[8], p_z = X[9], p_vx = X[10], p_vy = X[11], p_vz = X[12];
    double r = sqrt(x * x + y * y + z * z);
    double v = sqrt(vx * vx + vy * vy + vz * vz);
    double rdotv = x * vx + y * vy + z * vz;
    double pvdotv = p_vx * vx + p_vy * vy + p_vz * vz;
    double b = m[G_b], C = m[G_C], KD = m[G_KD], kr = m[G_kr];
    double g = 1 / r / r;
    double norm_pv = sqrt(p_vx * p_vx + p_vy * p_vy + p_vz * p_vz);
    double D = KD * exp(-kr * (r - 1));

    double p_xdot = -kr * KD / mass * v * exp(-kr * (r - 1)) * x / r * pvdotv + g * (p_vx * (1 - 3 * x * x / r / r) / r - p_vy * 3 * x * y / r / r / r - p_vz * 3 * x * z / r / r / r);
    double p_ydot = -kr * KD / mass * v * exp(-kr * (r - 1)) * y / r * pvdotv + g * (-p_vx * 3 * y * x / r / r / r + p_vy * (1 - 3 * y * y / r / r) / r - p_vz * 3 * y * z / r / r / r);
    double p_zdot = -kr * KD / mass * v * exp(-kr * (r - 1)) * z / r * pvdotv + g * (-p_vx * 3 * z * x / r / r / r - p_vy * 3 * z * y / r / r / r + p_vz * (1 - 3 * z * z / r / r) / r);
    double p_vxdot = -p_x + KD / mass * exp(-kr * (r - 1)) * (pvdotv * vx / v + p_vx * v);
    double p_vydot = -p_y + KD / mass * exp(-kr * (r - 1)) * (pvdotv * vy / v + p_vy * v);
    double p_vzdot = -p_z + KD / mass * exp(-kr * (r - 1)) * (pvdotv * vz / v + p_vz * v);

    double prdotdotpv = p_xdot * p_vx + p_ydot * p_vy + p_zdot * p_vz;
    double prdotpvdot = p_x * p_vxdot + p_y * p_vydot + p_z * p_vzdot;
    double prdotpv = p_x * p_vx + p_y * p_vy + p_z * p_vz;
    double pvdotdotv = p_vxdot * vx + p_vydot * vy + p_vzdot * vz;
    double pvdotdotpv = p_vxdot * p_vx + p_vydot * p_vy + p_vzdot * p_vz;
    double vdotg = vx * g * x / r + vy * g * y / r + vz * g * z / r;
    double pvdotg = p_vx * g * x / r + p_vy * g * y / r + p_vz * g * z / r;

    double au = 2 * norm_pv * C / mass * pvdotv
        + 2 * pvdotv * C / mass * norm_pv
        - b / mass * (2 * pvdotv * pvdotv + norm_pv * norm_pv * v * v)
        - b / D * v * prdotpv - C / D * prdotpv / v * pvdotv / norm_pv;

    double bu = -2 * norm_pv * norm_pv * (vdotg + D / mass * v * v * v) + 2 * v * v * pvdotdotpv
        - 2 * pvdotv * (pvdotg + D / mass * v * pvdotv - pvdotdotv)
        + b / C * (2 * norm_pv * pvdotv * (vdotg + D / mass * v * v * v) + norm_pv * v * v * (pvdotg + D / mass * v * pvdotv - pvdotdotv) - v * v * pvdotv / norm_pv * pvdotdotpv)
        - mass / D * kr * rdotv / r * v * prdotpv + mass / D * prdotpv / v * (vdotg + D / mass * v * v * v) - mass / D * v * (prdotdotpv + prdotpvdot);

    return bu / au;
}

/* goddard.cpp:104-185 */
static void goddard_control(const so_problem *p, double t, const double *X, double *u)
{
    const double *m = p->mparams;
    double mass = X[6], p_vx = X[10], p_vy = X[11], p_vz = X[12], p_mass = X[13];
    double b = m[G_b], C = m[G_C];
    double norm_pv = sqrt(p_vx * p_vx + p_vy * p_vy + p_vz * p_vz);
    double alpha_u = 0;
    double Switch = m[G_mu1] - b * p_mass - C / mass * norm_pv;

    if (m[G_mu2] > 0) {
        if (Switch < 0) alpha_u = -Switch / 2 / m[G_mu2];
        else alpha_u = 0;
    } else {
        if (t <= p->sw[0]) alpha_u = 1.0;
        else if (t > p->sw[0] && t <= p->sw[1]) {
            if (m[G_sing] < 0) alpha_u = goddard_singular(p, X);
            else alpha_u = m[G_sing];
        } else alpha_u = 0;
    }
    u[0] = -p_vx * alpha_u / norm_pv;
    u[1] = -p_vy * alpha_u / norm_pv;
    u[2] = -p_vz * alpha_u / norm_pv;
    double norm_u = fabs(alpha_u);
    double u_max = m[G_umax];
    if (norm_u > u_max) {
        u[0] = u[0] / norm_u * u_max;
        u[1] = u[1] / norm_u * u_max;
        u[2] = u[2] / norm_u * u_max;
    }
}

/* goddard.cpp:48-101 */
static void goddard_rhs(const so_problem *p, double t, const double *X, double *Xdot)
{
    const double *m = p->mparams;
    double x = X[0], y = X[1], z = X[2], vx = X[3], vy = X[4], vz = X[5], mass = X[6];
    double p_x = X[7], p_y = X[8], p_z = X[9], p_vx = X[10], p_vy = X[11], p_vz = X[12];
    double r = sqrt(x * x + y * y + z * z);
    double v = sqrt(vx * vx + vy * vy + vz * vz);
    double pvdotv = p_vx * vx + p_vy * vy + p_vz * vz;
    double b = m[G_b], C = m[G_C], KD = m[G_KD], kr = m[G_kr];
    double g = 1 / r / r;
    double u[3];
    goddard_control(p, t, X, u);
    double norm_u = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    double pvdotu = p_vx * u[0] + p_vy * u[1] + p_vz * u[2];

    Xdot[0] = vx;
    Xdot[1] = vy;
    Xdot[2] = vz;
    Xdot[3] = -KD * v * vx * exp(-kr * (r - 1)) / mass - g * x / r + C * u[0] / mass;
    Xdot[4] = -KD * v * vy * exp(-kr * (r - 1)) / mass - g * y / r + C * u[1] / mass;
    Xdot[5] = -KD * v * vz * exp(-kr * (r - 1)) / mass - g * z / r + C * u[2] / mass;
    Xdot[6] = -b * norm_u;
    Xdot[7] = -kr * KD / mass * v * exp(-kr * (r - 1)) * x / r * pvdotv + g * (p_vx * (1 - 3 * x * x / r / r) / r - p_vy * 3 * x * y / r / r / r - p_vz * 3 * x * z / r / r / r);
    Xdot[8] = -kr * KD / mass * v * exp(-kr * (r - 1)) * y / r * pvdotv + g * (-p_vx * 3 * y * x / r / r / r + p_vy * (1 - 3 * y * y / r / r) / r - p_vz * 3 * y * z / r / r / r);
    Xdot[9] = -kr * KD / mass * v * exp(-kr * (r - 1)) * z / r * pvdotv + g * (-p_vx * 3 * z * x / r / r / r - p_vy * 3 * z * y / r / r / r + p_vz * (1 - 3 * z * z / r / r) / r);
    Xdot[10] = -p_x + KD / mass * exp(-kr * (r - 1)) * (pvdotv * vx / v + p_vx * v);
    Xdot[11] = -p_y + KD / mass * exp(-kr * (r - 1)) * (pvdotv * vy / v + p_vy * v);
    Xdot[12] = -p_z + KD / mass * exp(-kr * (r - 1)) * (pvdotv * vz / v + p_vz * v);
    Xdot[13] = -KD * exp(-kr * (r - 1)) / mass / mass * v * pvdotv + C / mass / mass * pvdotu;
}

/* goddard.cpp:256-295 */
static double goddard_H(const so_problem *p, double t, const double *X)
{
    const double *m = p->mparams;
    double x = X[0], y = X[1], z = X[2], vx = X[3], vy = X[4], vz = X[5], mass = X[6];
    double p_x = X[7], p_y = X[8], p_z = X[9], p_vx = X[10], p_vy = X[11], p_vz = X[12], p_mass = X[13];
    double r = sqrt(x * x + y * y + z * z);
    double v = sqrt(vx * vx + vy * vy + vz * vz);
    double b = m[G_b], C = m[G_C], KD = m[G_KD], kr = m[G_kr];
    double g = 1 / r / r;
    double u[3];
    goddard_control(p, t, X, u);
    double norm_u = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    double H = m[G_mu1] * norm_u + m[G_mu2] * norm_u * norm_u
        + p_x * vx + p_y * vy + p_z * vz
        + p_vx * (-KD * v * vx * exp(-kr * (r - 1)) / mass - g * x / r + C * u[0] / mass)
        + p_vy * (-KD * v * vy * exp(-kr * (r - 1)) / mass - g * y / r + C * u[1] / mass)
        + p_vz * (-KD * v * vz * exp(-kr * (r - 1)) / mass - g * z / r + C * u[2] / mass)
        - p_mass * b * norm_u;
    return H;
}

/* =================== doubleIntegrator (src/models/doubleIntegrator) ====================== */

/* doubleIntegrator.cpp:218-259 */
static void di_control(const so_problem *p, const double *X, double *u)
{
    double a_max = p->mparams[D_amax], u_max = p->mparams[D_umax];
    u[0] = -X[9] / a_max;
    u[1] = -X[10] / a_max;
    u[2] = -X[11] / a_max;
    double norm_u = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    if (norm_u > u_max) {
        u[0] = u[0] / norm_u * u_max;
        u[1] = u[1] / norm_u * u_max;
        u[2] = u[2] / norm_u * u_max;
    }
}
/* doubleIntegrator.cpp:67-108 */
static void di_rhs(const so_problem *p, const double *X, double *Xdot)
{
    double a_max = p->mparams[D_amax];
    double u[3];
    di_control(p, X, u);
    Xdot[0] = X[3]; Xdot[1] = X[4]; Xdot[2] = X[5];
    Xdot[3] = a_max * u[0]; Xdot[4] = a_max * u[1]; Xdot[5] = a_max * u[2];
    Xdot[6] = 0; Xdot[7] = 0; Xdot[8] = 0;
    Xdot[9] = -X[6]; Xdot[10] = -X[7]; Xdot[11] = -X[8];
}
/* doubleIntegrator.cpp:264-300 (isJac == 0) */
static double di_H(const so_problem *p, const double *X)
{
    double a_max = p->mparams[D_amax];
    double u[3];
    di_control(p, X, u);
    double norm_u = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    return p->mparams[D_muT] + a_max * a_max * norm_u * norm_u / 2 + X[6] * X[3] + X[7] * X[4] + X[8] * X[5]
         + a_max * (X[9] * u[0] + X[10] * u[1] + X[11] * u[2]);
}

/* ========================= covid19 (src/models/covid19/covid19.cpp) ====================== */

/* covid19.cpp:98-129 */
static double covid_control(const so_problem *p, const double *X)
{
    const double *m = p->mparams;
    double S = X[0], I = X[2], pS = X[4], pE = X[5];
    double c = (pE - pS) * S * I / m[C_Tinf] / m[C_N] * m[C_R0];
    if (c <= m[C_umin]) c = m[C_umin];
    if (c >= m[C_umax]) c = m[C_umax];
    return c;
}
/* covid19.cpp:53-95 (note the costate rows use R = X[3], as the reference does) */
static void covid_rhs(const so_problem *p, const double *X, double *Xdot)
{
    const double *m = p->mparams;
    double S = X[0], E = X[1], I = X[2], R = X[3], pS = X[4], pE = X[5], pI = X[6], pR = X[7];
    double R0 = m[C_R0], Tinf = m[C_Tinf], Tinc = m[C_Tinc], N = m[C_N], Imax = m[C_Imax], muI = m[C_muI];
    double u = covid_control(p, X);
    double Rt = R0 * (1 - u);
    double Ipen = 0;
    if (I >= Imax) Ipen = -muI * (I - Imax);
    Xdot[0] = -Rt / Tinf / N * S * I;
    Xdot[1] = Rt / Tinf / N * S * I - E / Tinc;
    Xdot[2] = E / Tinc - I / Tinf;
    Xdot[3] = I / Tinf;
    Xdot[4] = (pS - pE) * R * I / Tinf / N;
    Xdot[5] = (pE - pI) / Tinc;
    Xdot[6] = (pS - pE) * R * S / Tinf / N + (pI - pR) / Tinf + Ipen;
    Xdot[7] = 0;
}
/* covid19.cpp:132-168 */
static double covid_H(const so_problem *p, const double *X)
{
    const double *m = p->mparams;
    double S = X[0], E = X[1], I = X[2], pS = X[4], pE = X[5], pI = X[6], pR = X[7];
    double R0 = m[C_R0], Tinf = m[C_Tinf], Tinc = m[C_Tinc], N = m[C_N], Imax = m[C_Imax], muI = m[C_muI];
    double u = covid_control(p, X);
    double Rt = R0 * (1 - u);
    double Ipen = 0;
    if (I >= Imax) Ipen = muI * (I - Imax) * (I - Imax) / 2;
    return u * u / 2 + Ipen
        + pS * (-Rt / Tinf / N * S * I)
        + pE * (Rt / Tinf / N * S * I - E / Tinc)
        + pI * (E / Tinc - I / Tinf)
        + pR * (I / Tinf);
}

/* =============== vtolUAV + obstacle map (src/models/vtolUAV, src/maps/obstacle) =========== */

/* obstacle.cpp:155-163 (Function), :168-178 (Gradient), :183-231, :236-316.
 * Waypoint penalties are commented out in the reference (obstacle.cpp:160,173): funcWP = 0. */
void so_obstacle_eval(const so_problem *p, const double *position, double *func, double *grad)
{
    const so_obstacles *o = p->obs;
    double muObs = p->mparams[V_muObs];
    double funcObs = 0, g0 = 0, g1 = 0, g2 = 0;
    int n = o ? o->n : 0;
    for (int i = 0; i < n; ++i) {
        double x = o->pos[i][0], y = o->pos[i][1], z = o->pos[i][2];
        double radx = o->rad[i][0], rady = o->rad[i][1], radz = o->rad[i][2];
        if (o->type[i] == 0) {
            double hx = position[0] - x, hy = position[1] - y, hz = position[2] - z;
            double d = sqrt(hx * hx + hy * hy + hz * hz);
            {   /* function: full ellipsoid radius (obstacle.cpp:203) */
                double rad = d / sqrt(hx * hx / radx / radx + hy * hy / rady / rady + hz * hz / radz / radz);
                double h = (d - rad) / muObs;
                funcObs = funcObs + (1 - tanh(h)) / 2;
            }
            {   /* gradient: z ignored in the radius (obstacle.cpp:260-265) */
                double rad = d / sqrt(hx * hx / radx / radx + hy * hy / rady / rady);
                double h = (d - rad) / muObs;
                double sq = sqrt(hx * hx / radx / radx + hy * hy / rady / rady);
                double rho2 = (radx * radx - rady * rady) / (radx * radx * rady * rady) / sq / sq / sq;
                g0 = g0 - hx / d * (1 - hy * hy * rho2) / muObs * (1 - tanh(h) * tanh(h)) / 2;
                g1 = g1 - hy / d * (1 + hx * hx * rho2) / muObs * (1 - tanh(h) * tanh(h)) / 2;
                g2 = g2 - 0;
            }
        } else if (o->type[i] == 1) {
            double hx = (fabs(position[0] - x) - radx) / muObs;
            double hy = (fabs(position[1] - y) - rady) / muObs;
            double hz = (fabs(position[2] - z) - radz) / muObs;
            funcObs = funcObs + (1 - tanh(hx)) * (1 - tanh(hy)) * (1 - tanh(hz)) / 8;
            g0 = g0 - (position[0] - x) / fabs(position[0] - x) / muObs * (1 - tanh(hx) * tanh(hx)) * (1 - tanh(hy)) * (1 - tanh(hz)) / 8;
            g1 = g1 - (position[1] - y) / fabs(position[1] - y) / muObs * (1 - tanh(hy) * tanh(hy)) * (1 - tanh(hx)) * (1 - tanh(hz)) / 8;
            g2 = g2 - (position[2] - z) / fabs(position[2] - z) / muObs * (1 - tanh(hz) * tanh(hz)) * (1 - tanh(hx)) * (1 - tanh(hy)) / 8;
        }
    }
    if (isnan(funcObs)) funcObs = 0.0;
    if (isnan(g0)) g0 = 0.0;
    if (isnan(g1)) g1 = 0.0;
    if (isnan(g2)) g2 = 0.0;
    double phi = p->mparams[V_phiObs], psi = p->mparams[V_psiWP];
    if (func) *func = phi * funcObs + psi * 0.0;
    if (grad) {
        grad[0] = phi * g0 + psi * 0.0;
        grad[1] = phi * g1 + psi * 0.0;
        grad[2] = phi * g2 + psi * 0.0;
    }
}

/* vtolUAV.cpp:110-148 */
static void vtol_control(const so_problem *p, const double *X, double *u)
{
    double a_max = p->mparams[V_amax], u_max = p->mparams[V_umax];
    u[0] = -X[9] / a_max;
    u[1] = -X[10] / a_max;
    u[2] = -X[11] / a_max;
    double norm_u = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    if (norm_u > u_max) {
        u[0] = u[0] / norm_u * u_max;
        u[1] = u[1] / norm_u * u_max;
        u[2] = u[2] / norm_u * u_max;
    }
}
/* vtolUAV.cpp:58-107 */
static void vtol_rhs(const so_problem *p, const double *X, double *Xdot)
{
    const double *m = p->mparams;
    double vx = X[3], vy = X[4], vz = X[5];
    double p_x = X[6], p_y = X[7], p_z = X[8], p_vx = X[9], p_vy = X[10], p_vz = X[11];
    double normV = sqrt(vx * vx + vy * vy + vz * vz);
    double a_max = m[V_amax], ca = m[V_ca], alphaV = m[V_alphaV], Vd = m[V_Vd];
    double u[3], grad[3];
    vtol_control(p, X, u);
    so_obstacle_eval(p, X, 0, grad);
    Xdot[0] = vx; Xdot[1] = vy; Xdot[2] = vz;
    Xdot[3] = a_max * u[0] - ca * vx * normV;
    Xdot[4] = a_max * u[1] - ca * vy * normV;
    Xdot[5] = a_max * u[2] - ca * vz * normV;
    Xdot[6] = 0 - grad[0];
    Xdot[7] = 0 - grad[1];
    Xdot[8] = 0 - grad[2];
    Xdot[9] = -p_x + ca * (p_vx * (normV + vx * vx / normV) + p_vy * (vy * vx / normV) + p_vz * (vz * vx / normV)) - alphaV * vx / normV * (normV - Vd);
    Xdot[10] = -p_y + ca * (p_vy * (normV + vy * vy / normV) + p_vx * (vx * vy / normV) + p_vz * (vz * vy / normV)) - alphaV * vy / normV * (normV - Vd);
    Xdot[11] = -p_z + ca * (p_vz * (normV + vz * vz / normV) + p_vx * (vx * vz / normV) + p_vy * (vy * vz / normV)) - alphaV * vz / normV * (normV - Vd);
}
/* vtolUAV.cpp:151-192 */
static double vtol_H(const so_problem *p, const double *X)
{
    const double *m = p->mparams;
    double vx = X[3], vy = X[4], vz = X[5];
    double p_x = X[6], p_y = X[7], p_z = X[8], p_vx = X[9], p_vy = X[10], p_vz = X[11];
    double normV = sqrt(vx * vx + vy * vy + vz * vz);
    double a_max = m[V_amax], ca = m[V_ca], alphaV = m[V_alphaV], Vd = m[V_Vd];
    double u[3], ObsValue = 0;
    vtol_control(p, X, u);
    double norm_u = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    so_obstacle_eval(p, X, &ObsValue, 0);
    return m[V_alphaT] * 1
        + alphaV / 2 * (normV - Vd) * (normV - Vd)
        + ObsValue
        + a_max * a_max * norm_u * norm_u / 2
        + p_x * vx + p_y * vy + p_z * vz
        + (p_vx * (a_max * u[0] - ca * vx * normV) + p_vy * (a_max * u[1] - ca * vy * normV) + p_vz * (a_max * u[2] - ca * vz * normV));
}

/* conditioning probe: uniform [-1,1] (tests only, see socp_oracle.h) */
static double noise_u(so_problem *p)
{
    p->noise_state = p->noise_state * 6364136223846793005ULL + 1442695040888963407ULL;
    return ((double)(p->noise_state >> 11) / 9007199254740992.0) * 2.0 - 1.0;
}
/* the argument of acos() in the chart conversions sits next to 1, where acos turns 1 ulp into
 * sqrt(ulp): the probe perturbs it like everything else */
static double noisy_unit(so_problem *p, double a)
{
    if (p && p->noise_ulps > 0) {
        a *= 1.0 + p->noise_ulps * 1.1102230246251565e-16 * noise_u(p);
        if (a > 1.0) a = 1.0;
        if (a < -1.0) a = -1.0;
    }
    return a;
}

/* ==================== interceptor (src/models/interceptor/interceptor.cpp) =============== */

/* interceptor.cpp:984-999 */
static double icp_mass(const so_problem *p, double t)
{
    const double *m = p->mparams;
    double qm = m[I_q] * m[I_mugft];
    double t1 = m[I_mprop] / m[I_q];
    if (p->stage == 1) return m[I_mempty] + m[I_mprop] - qm * t;
    return m[I_mempty] + m[I_mprop] - qm * t1;
}

typedef struct { double qm, mass, c_max, d, r, g, ft, eta, hr, alpha_max, u_max, muC; } icp_tmp;

/* the "temporary variables" block repeated at interceptor.cpp:293-306, :360-373, ... */
static void icp_common(const so_problem *p, double t, double h, icp_tmp *c)
{
    const double *m = p->mparams;
    c->qm = p->stage * m[I_q] * m[I_mugft];
    c->mass = icp_mass(p, t);
    c->c_max = m[I_c0] * exp(-h / m[I_hr]) * (m[I_mprop] + m[I_mempty]) / c->mass;
    c->d = m[I_d0] * exp(-h / m[I_hr]) * (m[I_mprop] + m[I_mempty]) / c->mass;
    c->r = h + I_R_EARTH;
    c->g = I_MU0 / c->r / c->r * m[I_mugft];
    c->ft = m[I_ve] * c->qm;
    c->eta = m[I_eta];
    c->hr = m[I_hr];
    c->alpha_max = m[I_alphamax];
    c->u_max = m[I_umax];
    c->muC = m[I_muC];
}

/* interceptor.cpp:342-389 */
static void icp_control_1(const so_problem *p, double t, const double *X, double *ctl)
{
    double h = X[0], v = X[1], gamma = X[2], p_v = X[7], p_gamma = X[8], p_chi = X[9];
    icp_tmp c;
    icp_common(p, t, h, &c);
    double c_max = c.c_max, ft = c.ft, mass = c.mass, alpha_max = c.alpha_max, eta = c.eta;
    double beta = atan2(p_chi, p_gamma * cos(gamma));
    double u = (p_gamma * (v * c_max * cos(beta) + ft * cos(beta) * alpha_max / mass / v)
                + p_chi * (v * c_max * sin(beta) / cos(gamma) + ft * sin(beta) / cos(gamma) * alpha_max / mass / v)
               ) / (p_v * (2 * eta * c_max * v * v + ft * alpha_max * alpha_max / mass) - c.muC);
    if (fabs(u) > c.u_max) u = c.u_max * u / fabs(u);
    ctl[0] = u;
    ctl[1] = beta;
}

/* interceptor.cpp:275-339 */
static void icp_rhs_1(const so_problem *p, double t, const double *X, double *Xdot)
{
    double h = X[0], v = X[1], gamma = X[2], chi = X[3], L = X[4];
    double p_h = X[6], p_v = X[7], p_gamma = X[8], p_chi = X[9], p_L = X[10], p_l = X[11];
    icp_tmp c;
    icp_common(p, t, h, &c);
    double mass = c.mass, c_max = c.c_max, d = c.d, r = c.r, g = c.g, ft = c.ft, eta = c.eta, hr = c.hr;
    double ctl[2];
    icp_control_1(p, t, X, ctl);
    double u = ctl[0], beta = ctl[1];
    double alpha = c.alpha_max * u;

    Xdot[0] = v * sin(gamma);
    Xdot[1] = -(d + eta * c_max * u * u) * v * v - g * sin(gamma) + ft * cos(alpha) / mass;
    Xdot[2] = v * c_max * u * cos(beta) - g / v * cos(gamma) + ft * sin(alpha) * cos(beta) / mass / v + v * cos(gamma) / r;
    Xdot[3] = v * c_max * u * sin(beta) / cos(gamma) + ft * sin(alpha) * sin(beta) / cos(gamma) / mass / v + v * cos(gamma) * tan(L) * sin(chi) / r;
    Xdot[4] = v * cos(gamma) * cos(chi) / r;
    Xdot[5] = v * cos(gamma) * sin(chi) / cos(L) / r;
    Xdot[6] = -p_v / hr * (d + eta * c_max * u * u) * v * v - 2 * g / r * (p_gamma / v * cos(gamma) + p_v * sin(gamma))
            + p_L * v * cos(gamma) * cos(chi) / r / r + p_gamma * v * cos(gamma) / r / r + p_gamma * v * c_max * u * cos(beta) / hr
            + p_l * v * cos(gamma) * sin(chi) / cos(L) / r / r + p_chi * v * cos(gamma) * tan(L) * sin(chi) / r / r + p_chi * v * c_max * u * sin(beta) / cos(gamma) / hr;
    Xdot[7] = -(p_L * cos(gamma) * cos(chi) / r + p_l * cos(gamma) * sin(chi) / cos(L) / r + p_h * sin(gamma)
            + p_gamma * (c_max * u * cos(beta) + g / v / v * cos(gamma) - ft * sin(alpha) * cos(beta) / mass / v / v + cos(gamma) / r)
            + p_chi * (c_max * u * sin(beta) / cos(gamma) - ft * sin(alpha) * sin(beta) / cos(gamma) / mass / v / v + cos(gamma) * tan(L) * sin(chi) / r)
            - p_v * 2 * (d + eta * c_max * u * u) * v);
    Xdot[8] = v * (p_L * sin(gamma) * cos(chi) / r + p_l * sin(gamma) * sin(chi) / cos(L) / r - p_h * cos(gamma))
            - g * (p_gamma / v * sin(gamma) - p_v * cos(gamma))
            + p_gamma * v * sin(gamma) / r + p_chi * v * sin(gamma) * tan(L) * sin(chi) / r
            - p_chi * (v * c_max * u * sin(beta) + ft * sin(alpha) * sin(beta) / mass / v) * sin(gamma) / cos(gamma) / cos(gamma);
    Xdot[9] = v * (p_L * cos(gamma) * sin(chi) / r - p_l * cos(gamma) * cos(chi) / cos(L) / r - p_chi * cos(gamma) * tan(L) * cos(chi) / r);
    Xdot[10] = -p_l * v * cos(gamma) * sin(chi) * sin(L) / cos(L) / cos(L) / r - p_chi * v * cos(gamma) * (1 + tan(L) * tan(L)) * sin(chi) / r;
    Xdot[11] = 0.0;
}

/* interceptor.cpp:392-442 */
static double icp_H_1(const so_problem *p, double t, const double *X)
{
    double h = X[0], v = X[1], gamma = X[2], chi = X[3], L = X[4];
    double p_h = X[6], p_v = X[7], p_gamma = X[8], p_chi = X[9], p_L = X[10], p_l = X[11];
    icp_tmp c;
    icp_common(p, t, h, &c);
    double mass = c.mass, c_max = c.c_max, d = c.d, r = c.r, g = c.g, ft = c.ft, eta = c.eta;
    double ctl[2];
    icp_control_1(p, t, X, ctl);
    double u = ctl[0], beta = ctl[1];
    double alpha = c.alpha_max * u;
    return p_L * v * cos(gamma) * cos(chi) / r
        + p_l * v * cos(gamma) * sin(chi) / cos(L) / r
        + p_h * v * sin(gamma)
        + p_gamma * (v * c_max * u * cos(beta) - g / v * cos(gamma) + ft * sin(alpha) * cos(beta) / mass / v + v * cos(gamma) / r)
        + p_chi * (v * c_max * u * sin(beta) / cos(gamma) + ft * sin(alpha) * sin(beta) / cos(gamma) / mass / v + v * cos(gamma) * tan(L) * sin(chi) / r)
        - p_v * ((d + eta * c_max * u * u) * v * v + g * sin(gamma) - ft * cos(alpha) / mass)
        + c.muC * u * u / 2;
}

/* interceptor.cpp:519-566 */
static void icp_control_2(const so_problem *p, double t, const double *X, double *ctl)
{
    double h = X[0], v = X[1], theta = X[2], p_v = X[7], p_theta = X[8], p_phi = X[9];
    icp_tmp c;
    icp_common(p, t, h, &c);
    double c_max = c.c_max, ft = c.ft, mass = c.mass, alpha_max = c.alpha_max, eta = c.eta;
    double beta = atan2(-p_phi, p_theta * cos(theta));
    double u = (p_theta * (v * c_max * cos(beta) + ft * cos(beta) * alpha_max / mass / v)
                - p_phi * (v * c_max * sin(beta) / cos(theta) + ft * sin(beta) / cos(theta) * alpha_max / mass / v)
               ) / (p_v * (2 * eta * c_max * v * v + ft * alpha_max * alpha_max / mass) - c.muC);
    if (fabs(u) > c.u_max) u = c.u_max * u / fabs(u);
    ctl[0] = u;
    ctl[1] = beta;
}

/* interceptor.cpp:445-516 */
static void icp_rhs_2(const so_problem *p, double t, const double *X, double *Xdot)
{
    double h = X[0], v = X[1], theta = X[2], phi = X[3], L = X[4];
    double p_h = X[6], p_v = X[7], p_theta = X[8], p_phi = X[9], p_L = X[10], p_l = X[11];
    icp_tmp c;
    icp_common(p, t, h, &c);
    double mass = c.mass, c_max = c.c_max, d = c.d, r = c.r, g = c.g, ft = c.ft, eta = c.eta, hr = c.hr;
    double ctl[2];
    icp_control_2(p, t, X, ctl);
    double u = ctl[0], beta = ctl[1];
    double alpha = c.alpha_max * u;

    Xdot[0] = -v * cos(theta) * cos(phi);
    Xdot[1] = -(d + eta * c_max * u * u) * v * v + g * cos(theta) * cos(phi) + ft * cos(alpha) / mass;
    Xdot[2] = v * c_max * u * cos(beta) + v * sin(theta) * (cos(phi) + sin(phi) * tan(L)) / r
            + (ft * sin(alpha) * cos(beta) / (mass * v) - g * sin(theta) * cos(phi) / v);
    Xdot[3] = -v * c_max * u * sin(beta) / cos(theta)
            + v * cos(theta) * (sin(phi) + tan(theta) * tan(theta) * (sin(phi) - tan(L) * cos(phi))) / r
            - (ft * sin(alpha) * sin(beta) / (mass * v * cos(theta)) + g * sin(phi) / (v * cos(theta)));
    Xdot[4] = v * cos(theta) * sin(phi) / r;
    Xdot[5] = v * sin(theta) / (r * cos(L));
    Xdot[6] = -p_v / hr * (d + eta * c_max * u * u) * v * v - 2 * g / r * (p_theta * sin(theta) * cos(phi) / v + p_phi * sin(phi) / cos(theta) / v - p_v * cos(theta) * cos(phi))
            + p_L * v * cos(theta) * sin(phi) / r / r + v * p_theta * sin(theta) * (cos(phi) + sin(phi) * tan(L)) / r / r + p_theta * v * c_max * u * cos(beta) / hr
            + p_l * v * sin(theta) / cos(L) / r / r + v * p_phi * cos(theta) * (sin(phi) + tan(theta) * tan(theta) * (sin(phi) - tan(L) * cos(phi))) / r / r - p_phi * v * c_max * u * sin(beta) / cos(theta) / hr;
    Xdot[7] = -(p_L * cos(theta) * sin(phi) / r + p_l * sin(theta) / (r * cos(L)) - p_h * cos(theta) * cos(phi)
            + p_theta * (c_max * u * cos(beta) + g / v / v * sin(theta) * cos(phi) - ft * sin(alpha) * cos(beta) / mass / v / v + sin(theta) * (cos(phi) + sin(phi) * tan(L)) / r)
            + p_phi * (-c_max * u * sin(beta) / cos(theta) + g / v / v * sin(phi) / cos(theta) + ft * sin(alpha) * sin(beta) / cos(theta) / mass / v / v + cos(theta) * (sin(phi) + tan(theta) * tan(theta) * (sin(phi) - tan(L) * cos(phi))) / r)
            - p_v * 2 * (d + eta * c_max * u * u) * v);
    Xdot[8] = -v * (-p_L * sin(theta) * sin(phi) / r + p_l * cos(theta) / (r * cos(L)) + p_h * sin(theta) * cos(phi))
            - g * (-p_theta * cos(theta) * cos(phi) / v - p_phi * sin(phi) * tan(theta) / (v * cos(theta)) - p_v * sin(theta) * cos(phi))
            - p_theta * v * cos(theta) * (cos(phi) + sin(phi) * tan(L)) / r + p_phi * v * sin(theta) * (sin(phi) + tan(theta) * tan(theta) * (sin(phi) - tan(L) * cos(phi))) / r
            - p_phi * v * cos(theta) * (2 * tan(theta) * (1 + tan(theta) * tan(theta)) * (sin(phi) - tan(L) * cos(phi))) / r
            - p_phi * (-v * c_max * u * sin(beta) - ft * sin(alpha) * sin(beta) / mass / v) * tan(theta) / cos(theta);
    Xdot[9] = -v * (p_h * cos(theta) * sin(phi) + p_L * cos(theta) * cos(phi) / r)
            - g * (p_theta * sin(theta) * sin(phi) / v - p_phi * cos(phi) / (v * cos(theta)) - p_v * cos(theta) * sin(phi))
            - p_theta * (v * sin(theta) * (-sin(phi) + cos(phi) * tan(L)) / r)
            - p_phi * v * cos(theta) * (cos(phi) + tan(theta) * tan(theta) * (cos(phi) + tan(L) * sin(phi))) / r;
    Xdot[10] = -p_l * v * sin(theta) * tan(L) / cos(L) / r - v * (1 + tan(L) * tan(L)) * (p_theta * sin(theta) * sin(phi) - p_phi * cos(theta) * cos(phi) * tan(theta) * tan(theta)) / r;
    Xdot[11] = 0.0;
}

/* interceptor.cpp:569-619 */
static double icp_H_2(const so_problem *p, double t, const double *X)
{
    double h = X[0], v = X[1], theta = X[2], phi = X[3], L = X[4];
    double p_h = X[6], p_v = X[7], p_theta = X[8], p_phi = X[9], p_L = X[10], p_l = X[11];
    icp_tmp c;
    icp_common(p, t, h, &c);
    double mass = c.mass, c_max = c.c_max, d = c.d, r = c.r, g = c.g, ft = c.ft, eta = c.eta;
    double ctl[2];
    icp_control_2(p, t, X, ctl);
    double u = ctl[0], beta = ctl[1];
    double alpha = c.alpha_max * u;
    return p_L * v * cos(theta) * sin(phi) / r
        + p_l * v * sin(theta) / (r * cos(L))
        - p_h * v * cos(theta) * cos(phi)
        + p_theta * (v * c_max * u * cos(beta) + v * sin(theta) * (cos(phi) + sin(phi) * tan(L)) / r + (ft * sin(alpha) * cos(beta) / (mass * v) - g * sin(theta) * cos(phi) / v))
        + p_phi * (-v * c_max * u * sin(beta) / cos(theta) + v * cos(theta) * (sin(phi) + tan(theta) * tan(theta) * (sin(phi) - tan(L) * cos(phi))) / r - (ft * sin(alpha) * sin(beta) / (mass * v * cos(theta)) + g * sin(phi) / (v * cos(theta))))
        - p_v * ((d + eta * c_max * u * u) * v * v - g * cos(theta) * cos(phi) - ft * cos(alpha) / mass)
        + c.muC * u * u / 2;
}

/* Jacobians of the Cartesian embedding wrt each chart (interceptor.cpp:651-699, 765-813) */
static void icp_jac_pos(double J[6][6], double L, double l, double r)
{
    J[0][0] = cos(L) * cos(l); J[1][0] = -r * sin(L) * cos(l); J[2][0] = -r * cos(L) * sin(l);
    J[3][0] = 0.0; J[4][0] = 0.0; J[5][0] = 0.0;
    J[0][1] = cos(L) * sin(l); J[1][1] = -r * sin(L) * sin(l); J[2][1] = r * cos(L) * cos(l);
    J[3][1] = 0.0; J[4][1] = 0.0; J[5][1] = 0.0;
    J[0][2] = sin(L); J[1][2] = r * cos(L); J[2][2] = 0.0; J[3][2] = 0.0; J[4][2] = 0.0; J[5][2] = 0.0;
}
static void icp_jac1(double J[6][6], double L, double l, double r, double v, double gamma, double chi)
{
    icp_jac_pos(J, L, l, r);
    J[0][3] = 0.0; J[1][3] = (-cos(L) * cos(l) * cos(gamma) * cos(chi) - sin(L) * cos(l) * sin(gamma)) * v;
    J[2][3] = (sin(L) * sin(l) * cos(gamma) * cos(chi) - cos(l) * cos(gamma) * sin(chi) - cos(L) * sin(l) * sin(gamma)) * v;
    J[3][3] = (sin(L) * cos(l) * sin(gamma) * cos(chi) + sin(l) * sin(gamma) * sin(chi) + cos(L) * cos(l) * cos(gamma)) * v;
    J[4][3] = (sin(L) * cos(l) * cos(gamma) * sin(chi) - sin(l) * cos(gamma) * cos(chi)) * v;
    J[5][3] = (-sin(L) * cos(l) * cos(gamma) * cos(chi) - sin(l) * cos(gamma) * sin(chi) + cos(L) * cos(l) * sin(gamma)) * v;
    J[0][4] = 0.0; J[1][4] = (-cos(L) * sin(l) * cos(gamma) * cos(chi) - sin(L) * sin(l) * sin(gamma)) * v;
    J[2][4] = (-sin(L) * cos(l) * cos(gamma) * cos(chi) - sin(l) * cos(gamma) * sin(chi) + cos(L) * cos(l) * sin(gamma)) * v;
    J[3][4] = (sin(L) * sin(l) * sin(gamma) * cos(chi) - cos(l) * sin(gamma) * sin(chi) + cos(L) * sin(l) * cos(gamma)) * v;
    J[4][4] = (sin(L) * sin(l) * cos(gamma) * sin(chi) + cos(l) * cos(gamma) * cos(chi)) * v;
    J[5][4] = (-sin(L) * sin(l) * cos(gamma) * cos(chi) + cos(l) * cos(gamma) * sin(chi) + cos(L) * sin(l) * sin(gamma)) * v;
    J[0][5] = 0.0; J[1][5] = (-sin(L) * cos(gamma) * cos(chi) + cos(L) * sin(gamma)) * v;
    J[2][5] = 0.0; J[3][5] = (-cos(L) * sin(gamma) * cos(chi) + sin(L) * cos(gamma)) * v;
    J[4][5] = -cos(L) * cos(gamma) * sin(chi) * v; J[5][5] = (cos(L) * cos(gamma) * cos(chi) + sin(L) * sin(gamma)) * v;
}
static void icp_jac2(double J[6][6], double L, double l, double r, double v, double theta, double phi)
{
    icp_jac_pos(J, L, l, r);
    J[0][3] = 0.0; J[1][3] = (-cos(L) * cos(l) * cos(theta) * sin(phi) + sin(L) * cos(l) * cos(theta) * cos(phi)) * v;
    J[2][3] = (sin(L) * sin(l) * cos(theta) * sin(phi) - cos(l) * sin(theta) + cos(L) * sin(l) * cos(theta) * cos(phi)) * v;
    J[3][3] = (sin(L) * cos(l) * sin(theta) * sin(phi) - sin(l) * cos(theta) + cos(L) * cos(l) * sin(theta) * cos(phi)) * v;
    J[4][3] = (-sin(L) * cos(l) * cos(theta) * cos(phi) + cos(L) * cos(l) * cos(theta) * sin(phi)) * v;
    J[5][3] = (-sin(L) * cos(l) * cos(theta) * sin(phi) - sin(l) * sin(theta) - cos(L) * cos(l) * cos(theta) * cos(phi)) * v;
    J[0][4] = 0.0; J[1][4] = (-cos(L) * sin(l) * cos(theta) * sin(phi) + sin(L) * sin(l) * cos(theta) * cos(phi)) * v;
    J[2][4] = (-sin(L) * cos(l) * cos(theta) * sin(phi) - sin(l) * sin(theta) - cos(L) * cos(l) * cos(theta) * cos(phi)) * v;
    J[3][4] = (sin(L) * sin(l) * sin(theta) * sin(phi) + cos(l) * cos(theta) + cos(L) * sin(l) * sin(theta) * cos(phi)) * v;
    J[4][4] = (-sin(L) * sin(l) * cos(theta) * cos(phi) + cos(L) * sin(l) * cos(theta) * sin(phi)) * v;
    J[5][4] = (-sin(L) * sin(l) * cos(theta) * sin(phi) + cos(l) * sin(theta) - cos(L) * sin(l) * cos(theta) * cos(phi)) * v;
    J[0][5] = 0.0; J[1][5] = (-sin(L) * cos(theta) * sin(phi) - cos(L) * cos(theta) * cos(phi)) * v;
    J[2][5] = 0.0; J[3][5] = (-cos(L) * sin(theta) * sin(phi) + sin(L) * sin(theta) * cos(phi)) * v;
    J[4][5] = (cos(L) * cos(theta) * cos(phi) + sin(L) * cos(theta) * sin(phi)) * v;
    J[5][5] = (cos(L) * cos(theta) * sin(phi) - sin(L) * cos(theta) * cos(phi)) * v;
}

/* 6x6 partial-pivot LU solve; stands in for Eigen's Matrix::lu().solve() (interceptor.cpp:715,829) */
static void lu6_solve(double A[6][6], const double *b, double *x)
{
    double lu[6][6];
    int piv[6];
    for (int i = 0; i < 6; ++i) { piv[i] = i; for (int j = 0; j < 6; ++j) lu[i][j] = A[i][j]; }
    for (int k = 0; k < 6; ++k) {
        int pr = k;
        double best = fabs(lu[k][k]);
        for (int i = k + 1; i < 6; ++i)
            if (fabs(lu[i][k]) > best) { best = fabs(lu[i][k]); pr = i; }
        if (pr != k) {
            for (int j = 0; j < 6; ++j) { double tmp = lu[k][j]; lu[k][j] = lu[pr][j]; lu[pr][j] = tmp; }
            int ti = piv[k]; piv[k] = piv[pr]; piv[pr] = ti;
        }
        for (int i = k + 1; i < 6; ++i) {
            lu[i][k] /= lu[k][k];
            for (int j = k + 1; j < 6; ++j) lu[i][j] -= lu[i][k] * lu[k][j];
        }
    }
    double y[6];
    for (int i = 0; i < 6; ++i) y[i] = b[piv[i]];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < i; ++j) y[i] -= lu[i][j] * y[j];
    for (int i = 5; i >= 0; --i) {
        for (int j = i + 1; j < 6; ++j) y[i] -= lu[i][j] * y[j];
        y[i] /= lu[i][i];
    }
    for (int i = 0; i < 6; ++i) x[i] = y[i];
}

/* costate transport: p_to = Jto * (Jfrom^-1 p_from), with the reference's component order
 * (interceptor.cpp:701-723) */
static void icp_costate_transport(double Jfrom[6][6], double Jto[6][6], const double *Xin, double *Xout)
{
    double pf[6] = {Xin[6], Xin[10], Xin[11], Xin[8], Xin[9], Xin[7]};
    double tmp[6], pt[6];
    lu6_solve(Jfrom, pf, tmp);
    for (int i = 0; i < 6; ++i) {
        double s = 0;
        for (int j = 0; j < 6; ++j) s += Jto[i][j] * tmp[j];
        pt[i] = s;
    }
    Xout[6] = pt[0]; Xout[7] = pt[5]; Xout[8] = pt[3]; Xout[9] = pt[4]; Xout[10] = pt[1]; Xout[11] = pt[2];
}

/* interceptor.cpp:622-727 */
static void icp_convert_12(so_problem *p, const double *X1, double *X2)
{
    const double eps = 1e-18;
    for (int i = 0; i < 12; ++i) X2[i] = X1[i];
    double h = X1[0], v = X1[1], gamma = X1[2], chi = X1[3], L = X1[4], l = X1[5];
    double r = h + I_R_EARTH;
    if (gamma == M_PI / 2.0) { X2[2] = 0; X2[3] = -M_PI; }
    else if (gamma == -M_PI / 2.0) { X2[2] = 0; X2[3] = 0; }
    else {
        double arg = noisy_unit(p, sqrt(sin(gamma) * sin(gamma) + cos(gamma) * cos(gamma) * cos(chi) * cos(chi)));
        X2[2] = acos(arg);
        if (cos(gamma) * sin(chi) < 0)
            X2[2] = -acos(arg);
        double sinPhi = cos(gamma) * cos(chi) / cos(X2[2]);
        if (fabs(sinPhi) < eps && sin(gamma) / cos(X2[2]) < 0) X2[3] = 0;
        else if (fabs(sinPhi) < eps && sin(gamma) / cos(X2[2]) > 0) X2[3] = -M_PI;
        else if (sinPhi > 0) X2[3] = acos(noisy_unit(p, -sin(gamma) / cos(X2[2])));
        else X2[3] = -acos(noisy_unit(p, -sin(gamma) / cos(X2[2])));
    }
    double J1[6][6], J2[6][6];
    icp_jac1(J1, L, l, r, v, gamma, chi);
    icp_jac2(J2, L, l, r, v, X2[2], X2[3]);
    icp_costate_transport(J1, J2, X1, X2);
}

/* interceptor.cpp:730-841 */
static void icp_convert_21(so_problem *p, const double *X2, double *X1)
{
    const double eps = 1e-18;
    for (int i = 0; i < 12; ++i) X1[i] = X2[i];
    double h = X2[0], v = X2[1], theta = X2[2], phi = X2[3], L = X2[4], l = X2[5];
    double r = h + I_R_EARTH;
    if (theta == M_PI / 2.0) { X1[2] = 0; X1[3] = M_PI / 2.0; }
    else if (theta == -M_PI / 2.0) { X1[2] = 0; X1[3] = -M_PI / 2.0; }
    else {
        double arg = noisy_unit(p, sqrt(sin(theta) * sin(theta) + cos(theta) * cos(theta) * sin(phi) * sin(phi)));
        X1[2] = acos(arg);
        if (cos(theta) * cos(phi) > 0)
            X1[2] = -acos(arg);
        double sinChi = sin(theta) / cos(X1[2]);
        if (fabs(sinChi) < eps && sin(phi) * cos(theta) / cos(X1[2]) > 0) X1[3] = 0;
        else if (fabs(sinChi) < eps && sin(phi) * cos(theta) / cos(X1[2]) < 0) X1[3] = -M_PI;
        else if (sinChi > 0) X1[3] = acos(noisy_unit(p, sin(phi) * cos(theta) / cos(X1[2])));
        else X1[3] = -acos(noisy_unit(p, sin(phi) * cos(theta) / cos(X1[2])));
    }
    double J1[6][6], J2[6][6];
    icp_jac1(J1, L, l, r, v, X1[2], X1[3]);
    icp_jac2(J2, L, l, r, v, theta, phi);
    icp_costate_transport(J2, J1, X2, X1);
}

/* interceptor.cpp:958-981 */
static void icp_set_chart(so_problem *p, double *X)
{
    if (fabs(cos(X[2])) >= I_CHART_LIMIT) return;
    double Y[12];
    if (p->chart == 1) { icp_convert_12(p, X, Y); p->chart = 2; }
    else { icp_convert_21(p, X, Y); p->chart = 1; }
    memcpy(X, Y, sizeof Y);
}

/* ================================ dispatch ============================================== */

static void so_rhs_exact(so_problem *p, double t, const double *X, double *dX);

void so_rhs(so_problem *p, double t, const double *X, double *dX)
{
    if (p->noise_ulps > 0) {
        /* backward-error model of a different libm / FMA contraction: the formulas are evaluated at
         * inputs moved by a few ulp (this is what carries rounding through cancellations such as
         * the singular control, goddard.cpp:188-253) and every output is moved by a few ulp */
        double Y[2 * SO_MAX_DIM];
        for (int i = 0; i < 2 * p->dim; ++i)
            Y[i] = X[i] * (1.0 + p->noise_ulps * 1.1102230246251565e-16 * noise_u(p));
        so_rhs_exact(p, t, Y, dX);
        for (int i = 0; i < 2 * p->dim; ++i)
            dX[i] *= 1.0 + p->noise_ulps * 1.1102230246251565e-16 * noise_u(p);
        return;
    }
    so_rhs_exact(p, t, X, dX);
}

static void so_rhs_exact(so_problem *p, double t, const double *X, double *dX)
{
    switch (p->model_id) {
    case SO_GODDARD: goddard_rhs(p, t, X, dX); break;
    case SO_DI: di_rhs(p, X, dX); break;
    case SO_COVID19: covid_rhs(p, X, dX); break;
    case SO_VTOL: vtol_rhs(p, X, dX); break;
    case SO_INTERCEPTOR:
        if (p->chart == 1) icp_rhs_1(p, t, X, dX); else icp_rhs_2(p, t, X, dX);
        break;
    }
}

int so_control(so_problem *p, double t, const double *X, double *u)
{
    switch (p->model_id) {
    case SO_GODDARD: goddard_control(p, t, X, u); return 3;
    case SO_DI: di_control(p, X, u); return 3;
    case SO_COVID19: u[0] = covid_control(p, X); return 1;
    case SO_VTOL: vtol_control(p, X, u); return 3;
    case SO_INTERCEPTOR:
        if (p->chart == 1) icp_control_1(p, t, X, u); else icp_control_2(p, t, X, u);
        return 2;
    }
    return 0;
}

static double so_hamiltonian_exact(so_problem *p, double t, const double *X);

double so_hamiltonian(so_problem *p, double t, const double *X)
{
    if (p->noise_ulps > 0) {           /* conditioning probe, same model as so_rhs */
        double Y[2 * SO_MAX_DIM];
        for (int i = 0; i < 2 * p->dim; ++i)
            Y[i] = X[i] * (1.0 + p->noise_ulps * 1.1102230246251565e-16 * noise_u(p));
        return so_hamiltonian_exact(p, t, Y) * (1.0 + p->noise_ulps * 1.1102230246251565e-16 * noise_u(p));
    }
    return so_hamiltonian_exact(p, t, X);
}

static double so_hamiltonian_exact(so_problem *p, double t, const double *X)
{
    switch (p->model_id) {
    case SO_GODDARD: return goddard_H(p, t, X);
    case SO_DI: return di_H(p, X);
    case SO_COVID19: return covid_H(p, X);
    case SO_VTOL: return vtol_H(p, X);
    case SO_INTERCEPTOR: return p->chart == 1 ? icp_H_1(p, t, X) : icp_H_2(p, t, X);
    }
    return 0;
}

/* ================================ integrator ============================================ */

/* odeTools.cpp:89-98 (and the function-pointer twin :80-87 used by the interceptor):
 * X <- X + (h/6) * (F1 + (F4 + 2*(F2 + F3))) */
void so_rk4_step(so_problem *p, double t, double *X, double h)
{
    int N = 2 * p->dim;
    double F1[SO_MAX_N], F2[SO_MAX_N], F3[SO_MAX_N], F4[SO_MAX_N], Y[SO_MAX_N];
    so_rhs(p, t, X, F1);
    for (int i = 0; i < N; ++i) Y[i] = X[i] + (h / 2.0) * F1[i];
    so_rhs(p, t + h / 2.0, Y, F2);
    for (int i = 0; i < N; ++i) Y[i] = X[i] + (h / 2.0) * F2[i];
    so_rhs(p, t + h / 2.0, Y, F3);
    for (int i = 0; i < N; ++i) Y[i] = X[i] + h * F3[i];
    so_rhs(p, t + h, Y, F4);
    for (int i = 0; i < N; ++i)
        X[i] = X[i] + (h / 6.0) * (F1[i] + (F4[i] + 2.0 * (F2[i] + F3[i])));
    ++p->rk4_steps;
}

/* odeTools.cpp:128-146 (non-Boost branch) */
void so_integrate(so_problem *p, double *X, double t0, double tf, double dt)
{
    double t = t0;
    while (t < (tf - dt / 2)) {
        if (t + dt > tf) so_rk4_step(p, t, X, tf - t);
        else so_rk4_step(p, t, X, dt);
        t += dt;
    }
}

/* ---- adaptive Dormand-Prince 5(4), the reference's -D_USE_BOOST build -----------------------------
 * odeTools.cpp:131-134 calls
 *     boost::numeric::odeint::integrate_adaptive(
 *         make_dense_output<runge_kutta_dopri5<odeVector>>(odeIntTol, odeIntTol), model, X, t0, tf, dt)
 * Boost.Odeint is a third-party dependency that is neither vendored in the reference nor present in
 * this container (unpinned system package, CMakeLists.txt:72-79), so this is a restatement of its
 * published algorithm (boost/numeric/odeint 1.7x: stepper/runge_kutta_dopri5.hpp,
 * stepper/controlled_runge_kutta.hpp, stepper/dense_output_runge_kutta.hpp,
 * integrate/detail/integrate_adaptive.hpp) and PARITY IS UNPINNED against the real library:
 *   - stages and operation order of runge_kutta_dopri5::do_step_impl (coefficients pre-multiplied by
 *     dt, sums left to right), FSAL derivative, error estimate with the dc coefficients;
 *   - default_error_checker: max_i |xerr_i| / (eps_abs + eps_rel (|x_i| + |dt| |dxdt_i|)) on the state
 *     and derivative at the START of the step;
 *   - default_step_adjuster: reject if err > 1 and dt *= max(0.9 err^(-1/3), 1/5); accept otherwise and,
 *     if err < 0.5, dt *= 0.9 max(err, 5^-5)^(-1/5);
 *   - integrate_adaptive for dense-output steppers: whole steps while t + dt <= tf (with Boost's
 *     epsilon-guarded comparisons), then the step is clamped to land on tf, until tf - t <= epsilon. */
static void dopri5_try(so_problem *p, double t, const double *in, const double *k1, double dt,
                       double *out, double *k7, double *xerr)
{
    const int N = 2 * p->dim;
    const double a2 = 1.0 / 5, a3 = 3.0 / 10, a4 = 4.0 / 5, a5 = 8.0 / 9;
    const double b21 = 1.0 / 5, b31 = 3.0 / 40, b32 = 9.0 / 40, b41 = 44.0 / 45, b42 = -56.0 / 15, b43 = 32.0 / 9,
                 b51 = 19372.0 / 6561, b52 = -25360.0 / 2187, b53 = 64448.0 / 6561, b54 = -212.0 / 729,
                 b61 = 9017.0 / 3168, b62 = -355.0 / 33, b63 = 46732.0 / 5247, b64 = 49.0 / 176, b65 = -5103.0 / 18656;
    const double c1 = 35.0 / 384, c3 = 500.0 / 1113, c4 = 125.0 / 192, c5 = -2187.0 / 6784, c6 = 11.0 / 84;
    const double dc1 = c1 - 5179.0 / 57600, dc3 = c3 - 7571.0 / 16695, dc4 = c4 - 393.0 / 640,
                 dc5 = c5 - (-92097.0 / 339200), dc6 = c6 - 187.0 / 2100, dc7 = -1.0 / 40;
    double k2[SO_MAX_N], k3[SO_MAX_N], k4[SO_MAX_N], k5[SO_MAX_N], k6[SO_MAX_N], y[SO_MAX_N];
    for (int i = 0; i < N; ++i) y[i] = 1.0 * in[i] + dt * b21 * k1[i];
    so_rhs(p, t + dt * a2, y, k2);
    for (int i = 0; i < N; ++i) y[i] = 1.0 * in[i] + dt * b31 * k1[i] + dt * b32 * k2[i];
    so_rhs(p, t + dt * a3, y, k3);
    for (int i = 0; i < N; ++i) y[i] = 1.0 * in[i] + dt * b41 * k1[i] + dt * b42 * k2[i] + dt * b43 * k3[i];
    so_rhs(p, t + dt * a4, y, k4);
    for (int i = 0; i < N; ++i)
        y[i] = 1.0 * in[i] + dt * b51 * k1[i] + dt * b52 * k2[i] + dt * b53 * k3[i] + dt * b54 * k4[i];
    so_rhs(p, t + dt * a5, y, k5);
    for (int i = 0; i < N; ++i)
        y[i] = 1.0 * in[i] + dt * b61 * k1[i] + dt * b62 * k2[i] + dt * b63 * k3[i] + dt * b64 * k4[i] + dt * b65 * k5[i];
    so_rhs(p, t + dt, y, k6);
    for (int i = 0; i < N; ++i)
        out[i] = 1.0 * in[i] + dt * c1 * k1[i] + dt * c3 * k3[i] + dt * c4 * k4[i] + dt * c5 * k5[i] + dt * c6 * k6[i];
    so_rhs(p, t + dt, out, k7);
    for (int i = 0; i < N; ++i)
        xerr[i] = dt * dc1 * k1[i] + dt * dc3 * k3[i] + dt * dc4 * k4[i] + dt * dc5 * k5[i] + dt * dc6 * k6[i] + dt * dc7 * k7[i];
}

void so_integrate_adaptive(so_problem *p, double *X, double t0, double tf, double dt, double tol)
{
    const int N = 2 * p->dim;
    const double eps = 2.220446049250313e-16;
    double t = t0, k1[SO_MAX_N], out[SO_MAX_N], k7[SO_MAX_N], xerr[SO_MAX_N];
    int have_deriv = 0;
    if (!(dt > 0)) return;                             /* the reference only integrates forward */
    while (tf - t > eps) {                             /* less_with_sign(t, tf, dt) */
        while ((t + dt) - tf <= eps) {                 /* less_eq_with_sign(t + dt, tf, dt) */
            /* dense_output_runge_kutta::do_step: retry until the controller accepts (at most 500 times) */
            if (!have_deriv) { so_rhs(p, t, X, k1); have_deriv = 1; }
            int failed = 0, ok = 0;
            while (!ok && failed < 500) {
                dopri5_try(p, t, X, k1, dt, out, k7, xerr);
                double err = 0;
                for (int i = 0; i < N; ++i) {
                    double e = fabs(xerr[i]) / (tol + tol * (1.0 * fabs(X[i]) + fabs(dt) * fabs(k1[i])));
                    if (e > err || e != e) err = e;
                }
                if (err > 1.0 || err != err) {
                    double f = 0.9 * pow(err, -1.0 / 3.0);
                    dt *= (f > 0.2 && f == f) ? f : 0.2;
                    ++failed;
                    ++p->dopri_rejected;
                } else {
                    t += dt;
                    if (err < 0.5) {
                        double e5 = pow(5.0, -5.0);
                        if (err < e5) err = e5;
                        dt *= 0.9 * pow(err, -1.0 / 5.0);
                    }
                    ok = 1;
                }
            }
            if (!ok) return;                           /* Boost throws step_adjustment_error here */
            for (int i = 0; i < N; ++i) { X[i] = out[i]; k1[i] = k7[i]; }
            ++p->dopri_steps;
        }
        /* st.initialize(current_state, current_time, tf - t): the derivative is evaluated again */
        dt = tf - t;
        have_deriv = 0;
    }
}

/* interceptor.cpp:104-130 */
static void icp_model_int(so_problem *p, double t0, double *X, double tf)
{
    double t = t0;
    double dt = (tf - t0) / p->step_nbr;
    for (int i = 0; i < p->step_nbr; ++i) {
        icp_set_chart(p, X);
        so_rk4_step(p, t, X, dt);
        t += dt;
    }
}

/* model::ComputeTraj (model.hpp:77) -> ModelInt (model.hpp:395; goddard.cpp:298; covid19.cpp:171;
 * vtolUAV.cpp:195), and interceptor::ComputeTraj (interceptor.cpp:165-220) */
void so_traj(so_problem *p, double t0, const double *X0, double tf, double *Xf)
{
    int N = 2 * p->dim;
    double X[SO_MAX_N];
    for (int i = 0; i < N; ++i) X[i] = X0[i];
    if (p->model_id != SO_INTERCEPTOR) {
        double dt = (tf - t0) / p->step_nbr;
        if (p->integrator == 1) so_integrate_adaptive(p, X, t0, tf, dt, p->ode_tol);
        else so_integrate(p, X, t0, tf, dt);
    } else {
        p->chart = 1;
        double t1 = p->mparams[I_mprop] / p->mparams[I_q];
        if (t0 < t1) {
            p->stage = 1;
            if (tf > t1) {
                icp_model_int(p, t0, X, t1);
                p->stage = 0;
                icp_model_int(p, t1, X, tf);
            } else {
                icp_model_int(p, t0, X, tf);
            }
        } else {
            p->stage = 0;
            icp_model_int(p, t0, X, tf);
        }
        if (p->chart == 2) {
            double Y[12];
            icp_convert_21(p, X, Y);
            memcpy(X, Y, sizeof Y);
        }
    }
    for (int i = 0; i < N; ++i) Xf[i] = X[i];
}

/* ================================ residual ============================================== */

/* shooting.cpp:1579-1617 */
void so_timeline(so_problem *p, const double *x, double *tl)
{
    int M = p->num_multi;
    int nbr = 2 * p->dim * M;
    int cur = 0;
    int nsw = 0;
    for (int j = 0; j <= M; ++j) {
        if (p->mode_t[j] == SO_FIXED) {
            tl[j] = p->time[j];
            for (int k = cur + 1; k < j; ++k)
                tl[k] = tl[cur] + (k - cur) * (tl[j] - tl[cur]) / (j - cur);
            cur = j;
        }
        if (p->mode_t[j] == SO_FREE) {
            nbr += 1;
            tl[j] = x[nbr - 1];
            if (j < M) p->sw[nsw++] = tl[j];
            for (int k = cur + 1; k < j; ++k)
                tl[k] = tl[cur] + (k - cur) * (tl[j] - tl[cur]) / (j - cur);
            cur = j;
        }
    }
    /* SwitchingTimesUpdate (goddard.cpp:373, vtolUAV.cpp:266): the model keeps the list; a
     * shorter list leaves the old tail in place in practice (std::vector::resize) */
    if (p->model_id == SO_GODDARD || p->model_id == SO_VTOL) p->nsw = nsw;
}

/* boundary residual at the first node: model::InitialFunction (model.hpp:196-213) */
static void initial_function(const so_problem *p, const double *X, double *f)
{
    int n = p->dim;
    for (int j = 0; j < n; ++j)
        f[j] = (p->mode_X[0][j] == SO_FREE) ? X[j + n] : X[j] - p->Xb[0][j];
}

/* boundary residual at the last node: model::FinalFunction (model.hpp:90-103) and the overrides
 * interceptor.cpp:223-245, vtolUAV.cpp:223-241 */
static void final_function(const so_problem *p, const double *X, double *f)
{
    int n = p->dim, M = p->num_multi;
    const double *Xf = p->Xb[M];
    const int *mode = p->mode_X[M];
    const double *m = p->mparams;
    for (int j = 0; j < n; ++j) {
        if (mode[j] == SO_FREE) {
            f[j] = X[j + n];
            if (p->model_id == SO_INTERCEPTOR && j == 1) f[j] = X[j + n] + m[I_muV];
            if (p->model_id == SO_VTOL && j < 6) {
                int nWP_tot = (int)m[V_nWPtot], nWP = (int)m[V_nWP];
                f[j] = X[j + n] - m[V_invSigma] * (nWP_tot - nWP) * (X[j] - Xf[j]) - 0.02 * (X[j] - Xf[j]);
            }
        } else {
            f[j] = X[j] - Xf[j];
            if (p->model_id == SO_INTERCEPTOR) {
                if (j == 0) f[j] = f[j] / m[I_hr];
                if (j == 3 && fabs(cos(Xf[2])) < 1e-5) f[j] = X[j + n];
            }
        }
    }
}

/* shooting.cpp:1511-1576 with isJac == 0 (the two Model() calls there are dead code) */
static void multiple_shooting_function(const so_problem *p, int node, const double *X, const double *Xp, double *f)
{
    int n = p->dim;
    const double *Xd = p->Xb[node];
    for (int j = 0; j < n; ++j) {
        switch (p->mode_X[node][j]) {
        case SO_FIXED:
            f[j] = X[j] - Xd[j];
            f[j + n] = Xp[j] - Xd[j];
            break;
        case SO_FREE:
            /* model::SwitchingStateFunction: empty by default (model.hpp:339); vtolUAV.cpp:273-283 */
            if (p->model_id == SO_VTOL) {
                f[j] = X[j] - Xp[j];
                f[j + n] = (X[j + n] - Xp[j + n]);
                if (j < 6) f[j + n] = (X[j + n] - Xp[j + n]) - p->mparams[V_invSigma] * (X[j] - Xd[j]);
            }
            break;
        default:
            f[j] = X[j] - Xp[j];
            f[j + n] = X[j + n] - Xp[j + n];
            break;
        }
    }
}

/* shooting::ShootingFunction (shooting.cpp:918-993) */
void so_residual(so_problem *p, const double *x, double *fvec)
{
    int n = p->dim, N = 2 * n, M = p->num_multi;
    double tl[SO_MAX_NODES];
    so_timeline(p, x, tl);
    double X1[SO_MAX_N], Xtf[SO_MAX_N], f[2 * SO_MAX_N];
    for (int k = 0; k < N; ++k) X1[k] = x[k];
    int nbr = N * M;
    for (int k = 0; k < so_num_param(p); ++k) fvec[k] = 0.0;
    for (int i = 0; i < M; ++i) {
        double t1 = tl[i], t2 = tl[i + 1];
        so_traj(p, t1, X1, t2, Xtf);
        int index = N * (i + 1);
        if (i == 0) {
            initial_function(p, X1, f);
            for (int k = 0; k < n; ++k) fvec[k] = f[k];
            if (p->mode_t[0] != SO_FIXED) {           /* InitialHFunction, model.hpp:239-255 */
                fvec[N * M] = so_hamiltonian(p, tl[0], X1);
                nbr += 1;
            }
        }
        if (i < M - 1) {
            const double *Xp = &x[index];
            if (p->mode_t[i + 1] == SO_FREE) {
                /* SwitchingTimesFunction: H(X)-H(Xp) (model.hpp:299-305); goddard returns H(X)
                 * only (goddard.cpp:343-370) */
                if (p->model_id == SO_GODDARD) fvec[nbr] = so_hamiltonian(p, t2, Xtf);
                else fvec[nbr] = so_hamiltonian(p, t2, Xtf) - so_hamiltonian(p, t2, Xp);
                nbr += 1;
            }
            for (int k = 0; k < N; ++k) f[k] = 0.0;
            multiple_shooting_function(p, i + 1, Xtf, Xp, f);
            for (int k = 0; k < N; ++k) fvec[index + k] = f[k];
            for (int k = 0; k < N; ++k) X1[k] = Xp[k];
        }
        if (i == M - 1) {
            final_function(p, Xtf, f);
            for (int k = 0; k < n; ++k) fvec[k + n] = f[k];
            if (p->mode_t[M] != SO_FIXED) {
                /* FinalHFunction: model.hpp:133-147; interceptor adds muT (interceptor.cpp:270) */
                double H = so_hamiltonian(p, t2, Xtf);
                if (p->model_id == SO_INTERCEPTOR) H += p->mparams[I_muT];
                fvec[nbr] = H;
                nbr += 1;
            }
        }
    }
}

static int residual_cb(void *ud, int n, const double *x, double *fvec, int iflag)
{
    (void)n; (void)iflag;
    so_residual((so_problem *)ud, x, fvec);
    return 0;
}

/* ================= analytic Jacobian path (modelOrder == 1, hybrj) ======================== */
/* Only the double integrator implements it in the reference (doubleIntegrator.cpp:113-213, :264-300).
 * Extended state: X[0..N) then the sensitivity rows X[N(k+1) + i] = dX_k / dX0_i (shooting.cpp:1003). */
#define SO_NV (SO_MAX_N * (SO_MAX_N + 1))

int so_has_variational(const so_problem *p) { return p->model_id == SO_DI; }

/* doubleIntegrator::ModelJacobian: [f ; (df/dX) Phi] with the reference's CONSTANT df/dX (the velocity
 * rows carry -1 on the p_v columns whatever a_max and the saturation, "a revoir en cas de saturation",
 * doubleIntegrator.cpp:167-169) and its dense triple loop (zero terms included, same summation order) */
static void di_rhs_var(so_problem *p, const double *X, double *dX)
{
    const int N = 12;
    static const int col[12] = {3, 4, 5, 9, 10, 11, -1, -1, -1, 6, 7, 8};
    static const double val[12] = {1, 1, 1, -1, -1, -1, 0, 0, 0, -1, -1, -1};
    di_rhs(p, X, dX);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double acc = 0;
            for (int k = 0; k < N; ++k) acc += ((k == col[i]) ? val[i] : 0.0) * X[N * (k + 1) + j];
            dX[N + N * i + j] = acc;
        }
}

/* doubleIntegrator::Hamiltonian(isJac = 1): (dH/dX, dH/dt) */
static void di_H_grad(const double *X, double *g)
{
    g[0] = 0; g[1] = 0; g[2] = 0; g[3] = X[6]; g[4] = X[7]; g[5] = X[8];
    g[6] = X[3]; g[7] = X[4]; g[8] = X[5]; g[9] = -X[9]; g[10] = -X[10]; g[11] = -X[11]; g[12] = 0;
}

/* model::ComputeTraj(isJac = 1): the same fixed-step RK4 (odeTools.cpp:89-98, :128-146) on the
 * (2n+1) 2n vector */
void so_traj_var(so_problem *p, double t0, const double *X0, double tf, double *Xf)
{
    const int N = 2 * p->dim, NV = N * (N + 1);
    double X[SO_NV], F1[SO_NV], F2[SO_NV], F3[SO_NV], F4[SO_NV], Y[SO_NV];
    memcpy(X, X0, sizeof(double) * NV);
    const double dt = (tf - t0) / p->step_nbr;
    double t = t0;
    while (t < (tf - dt / 2)) {
        const double h = (t + dt > tf) ? tf - t : dt;
        di_rhs_var(p, X, F1);
        for (int i = 0; i < NV; ++i) Y[i] = X[i] + (h / 2.0) * F1[i];
        di_rhs_var(p, Y, F2);
        for (int i = 0; i < NV; ++i) Y[i] = X[i] + (h / 2.0) * F2[i];
        di_rhs_var(p, Y, F3);
        for (int i = 0; i < NV; ++i) Y[i] = X[i] + h * F3[i];
        di_rhs_var(p, Y, F4);
        for (int i = 0; i < NV; ++i) X[i] = X[i] + (h / 6.0) * (F1[i] + (F4[i] + 2.0 * (F2[i] + F3[i])));
        t += dt;
        ++p->rk4_steps;
    }
    memcpy(Xf, X, sizeof(double) * NV);
}

static void seed_identity(const double *state, int N, double *X)
{
    memset(X, 0, sizeof(double) * N * (N + 1));
    for (int i = 0; i < N; ++i) { X[i] = state[i]; X[N * (i + 1) + i] = 1; }
}

/* rows of d(boundary function)/d(node unknowns [, time]) as model::Initial/FinalFunction and the H
 * variants build them (model.hpp:104-120, :149-183, :214-226, :256-290); `width` = N or N + 1 */
static void boundary_jac(so_problem *p, double t, const double *X, const int *mode_X, int with_H, double *func)
{
    const int n = p->dim, N = 2 * n, width = with_H ? N + 1 : N;
    double f[SO_MAX_N], g[SO_MAX_N + 1];
    if (with_H) so_rhs(p, t, X, f);
    for (int j = 0; j < n; ++j) {
        const int row = (mode_X[j] == SO_FREE) ? j + n : j;
        for (int i = 0; i < N; ++i) func[width * j + i] = X[N * (row + 1) + i];
        if (with_H) func[width * j + N] = f[row];
    }
    if (with_H) {
        di_H_grad(X, g);
        for (int i = 0; i < N; ++i) {
            double acc = 0;
            for (int k = 0; k < N; ++k) acc += g[k] * X[N * (k + 1) + i];
            func[width * n + i] = acc;
        }
        double acc = 0;
        for (int k = 0; k < N; ++k) acc += g[k] * f[k];
        func[width * n + N] = acc + g[N];
    }
}

/* shooting::ShootingFunctionJacobian (shooting.cpp:996-1130) + the transpose of
 * StaticShootingFunctionJacobian (:889-893): fjac column-major, fjac[i + P j] = dF_i / dx_j */
int so_jacobian(so_problem *p, const double *x, double *fjac)
{
    if (!so_has_variational(p)) return -1;
    const int n = p->dim, N = 2 * n, M = p->num_multi, P = so_num_param(p), W4 = 4 * n + 1;
    double tl[SO_MAX_NODES];
    so_timeline(p, x, tl);
    double *J = (double *)calloc((size_t)P * P, sizeof(double));      /* row-major, as the reference fills it */
    double X_t0[SO_NV], X1[SO_NV], X_tf[SO_NV], Xp[SO_NV], func[(SO_MAX_DIM + 1) * (SO_MAX_N + 1)];
    double funcMS[SO_MAX_N * (4 * SO_MAX_DIM + 1)], sf[4 * SO_MAX_DIM + 1];
    seed_identity(x, N, X_t0);
    memcpy(X1, X_t0, sizeof X1);
    int nbr = N * M;
    for (int i = 0; i < M; ++i) {
        const double t1 = tl[i], t2 = tl[i + 1];
        so_traj_var(p, t1, X1, t2, X_tf);
        const int index = N * (i + 1);
        if (i == 0) {
            if (p->mode_t[0] == SO_FIXED) {
                boundary_jac(p, tl[0], X1, p->mode_X[0], 0, func);
                for (int k = 0; k < n; ++k)
                    for (int j = 0; j < N; ++j) J[P * k + j] = func[N * k + j];
            } else {
                boundary_jac(p, tl[0], X1, p->mode_X[0], 1, func);
                for (int k = 0; k < n; ++k) {
                    for (int j = 0; j < N; ++j) J[P * k + j] = func[(N + 1) * k + j];
                    J[P * k + nbr] = func[(N + 1) * k + N];
                }
                for (int j = 0; j < N; ++j) J[P * nbr + j] = func[(N + 1) * n + j];
                J[P * nbr + nbr] = func[(N + 1) * n + N];
                nbr += 1;
            }
        }
        if (i < M - 1) {
            seed_identity(x + index, N, Xp);
            /* shooting::MultipleShootingFunction(isJac = 1), shooting.cpp:1511-1576 */
            double fxt[SO_MAX_N], fxp[SO_MAX_N];
            so_rhs(p, t2, X_tf, fxt);
            so_rhs(p, t2, Xp, fxp);
            memset(funcMS, 0, sizeof funcMS);
            const int free_t = p->mode_t[i + 1] == SO_FREE;
            for (int j = 0; j < n; ++j) {
                const int m = p->mode_X[i + 1][j];
                if (m == SO_FIXED) {
                    for (int q = 0; q < N; ++q) {
                        funcMS[W4 * j + q] = X_tf[N * (j + 1) + q];
                        funcMS[W4 * (j + n) + N + q] = Xp[N * (j + 1) + q];
                    }
                    if (free_t) { funcMS[W4 * j + 4 * n] = fxt[j]; funcMS[W4 * (j + n) + 4 * n] = fxp[j]; }
                } else if (m == SO_CONTINUOUS) {
                    for (int q = 0; q < N; ++q) {
                        funcMS[W4 * j + q] = X_tf[N * (j + 1) + q];
                        funcMS[W4 * j + N + q] = -Xp[N * (j + 1) + q];
                        funcMS[W4 * (j + n) + q] = X_tf[N * (j + n + 1) + q];
                        funcMS[W4 * (j + n) + N + q] = -Xp[N * (j + n + 1) + q];
                    }
                    if (free_t) {
                        funcMS[W4 * j + 4 * n] = fxt[j] - fxp[j];
                        funcMS[W4 * (j + n) + 4 * n] = fxt[j + n] - fxp[j + n];
                    }
                }   /* FREE interior state: model::SwitchingStateFunction is empty by default (model.hpp:339) */
            }
            const int ncols = free_t ? W4 : 4 * n;
            for (int k = 0; k < N; ++k) {
                for (int j = 0; j < ncols; ++j) J[P * (index + k) + (index - N + j)] = funcMS[W4 * k + j];
                if (free_t) J[P * (index + k) + nbr] = funcMS[W4 * k + 4 * n];
            }
            if (free_t) {
                /* model::SwitchingTimesFunction(isJac = 1), model.hpp:306-327 */
                double gX[SO_MAX_N + 1], gP[SO_MAX_N + 1];
                di_H_grad(X_tf, gX);
                di_H_grad(Xp, gP);
                memset(sf, 0, sizeof sf);
                for (int q = 0; q < N; ++q)
                    for (int k = 0; k < N; ++k) {
                        sf[q] += gX[k] * X_tf[N * (k + 1) + q];
                        sf[N + q] -= gP[k] * Xp[N * (k + 1) + q];
                    }
                for (int k = 0; k < N; ++k) sf[4 * n] += gX[k] * fxt[k] - gP[k] * fxp[k];
                sf[4 * n] += gX[N] - gP[N];
                for (int j = 0; j < 4 * n; ++j) J[P * nbr + (index - N + j)] = sf[j];
                J[P * nbr + nbr] = sf[4 * n];
                nbr += 1;
            }
            memcpy(X1, Xp, sizeof X1);
        }
        if (i == M - 1) {
            if (p->mode_t[M] == SO_FIXED) {
                boundary_jac(p, t2, X_tf, p->mode_X[M], 0, func);
                for (int k = 0; k < n; ++k)
                    for (int j = 0; j < N; ++j) J[P * (n + k) + (N * i + j)] = func[N * k + j];
            } else {
                boundary_jac(p, t2, X_tf, p->mode_X[M], 1, func);
                for (int k = 0; k < n; ++k) {
                    for (int j = 0; j < N; ++j) J[P * (n + k) + (N * i + j)] = func[(N + 1) * k + j];
                    J[P * (n + k) + nbr] = func[(N + 1) * k + N];
                }
                for (int j = 0; j < N; ++j) J[P * nbr + (N * i + j)] = func[(N + 1) * n + j];
                J[P * nbr + nbr] = func[(N + 1) * n + N];
                nbr += 1;
            }
        }
    }
    for (int k = 0; k < P; ++k)
        for (int j = 0; j < P; ++j) fjac[k + P * j] = J[P * k + j];
    free(J);
    return 0;
}

static int residual_jac_cb(void *ud, int n, const double *x, double *fvec, double *fjac, int ldfjac, int iflag)
{
    (void)n; (void)ldfjac;
    if (iflag == 1) so_residual((so_problem *)ud, x, fvec);
    else so_jacobian((so_problem *)ud, x, fjac);
    return 0;
}

/* shooting::SolveShootingFunction, modelOrder == 1 branch (shooting.cpp:830-851) */
int so_solve_hybrj(so_problem *p, double *x, double xtol, int maxfev, int *nfev, int *njev, double *fnorm)
{
    int P = so_num_param(p);
    int lr = P * (P + 1) / 2;
    double *w = (double *)calloc((size_t)(P * P + lr + 7 * P), sizeof(double));
    double *fjac = w, *r = fjac + P * P, *qtf = r + lr, *fvec = qtf + P, *diag = fvec + P;
    double *wa1 = diag + P, *wa2 = wa1 + P, *wa3 = wa2 + P, *wa4 = wa3 + P;
    for (int i = 0; i < P; ++i) diag[i] = 1;
    int info = hybrj(residual_jac_cb, p, P, x, fvec, fjac, P, xtol, maxfev, diag, 1, 1.0, 0, nfev, njev,
                     r, lr, qtf, wa1, wa2, wa3, wa4);
    if (fnorm) *fnorm = mp_enorm(P, fvec);
    free(w);
    return info;
}

/* forward-difference Jacobian with MINPACK's step rule (fdjac1, dense), column-major */
void so_fdjac(so_problem *p, const double *x0, double epsfcn, double *fjac)
{
    int P = so_num_param(p);
    double *x = (double *)malloc(sizeof(double) * 3 * P), *f0 = x + P, *f1 = x + 2 * P;
    memcpy(x, x0, sizeof(double) * P);
    so_residual(p, x, f0);
    double eps = sqrt(epsfcn > 2.220446049250313e-16 ? epsfcn : 2.220446049250313e-16);
    for (int j = 0; j < P; ++j) {
        double temp = x[j], h = eps * fabs(temp);
        if (h == 0.) h = eps;
        x[j] = temp + h;
        so_residual(p, x, f1);
        x[j] = temp;
        for (int i = 0; i < P; ++i) fjac[i + j * P] = (f1[i] - f0[i]) / h;
    }
    free(x);
}

/* shooting::SolveShootingFunction (shooting.cpp:781-856), modelOrder == 0 branch, with the
 * solver constants of the shooting ctor (shooting.cpp:95-101) */
int so_solve(so_problem *p, double *x, double xtol, int maxfev, int *nfev, double *fnorm)
{
    int P = so_num_param(p);
    int lr = P * (P + 1) / 2;
    double *w = (double *)calloc((size_t)(P * P + lr + 7 * P), sizeof(double));
    double *fjac = w, *r = fjac + P * P, *qtf = r + lr, *fvec = qtf + P, *diag = fvec + P;
    double *wa1 = diag + P, *wa2 = wa1 + P, *wa3 = wa2 + P, *wa4 = wa3 + P;
    for (int i = 0; i < P; ++i) diag[i] = 1;
    int info = hybrd(residual_cb, p, P, x, fvec, xtol, maxfev, P - 1, P - 1, 1e-15, diag, 1, 1.0, 0,
                     nfev, fjac, P, r, lr, qtf, wa1, wa2, wa3, wa4);
    if (fnorm) *fnorm = mp_enorm(P, fvec);
    free(w);
    return info;
}

/* shooting.cpp:695-778 */
int so_continuation_param(so_problem *p, double *tab_param, double xtol, int maxfev, double step,
                          int param_idx, double goal, double step_min, int *calls)
{
    int P = so_num_param(p);
    double *Rdata = &p->mparams[param_idx];
    double Rstart = *Rdata;
    double bStep = step;
    double b = step < 1.0 ? step : 1.0;
    double b_prec = 0;
    *Rdata = (1 - b) * Rstart + b * goal;
    double *tmp = (double *)malloc(sizeof(double) * P);
    memcpy(tmp, tab_param, sizeof(double) * P);
    int condition = 1, ret = 0, nfev = 0;
    calls[0] = calls[1] = 0;
    while (condition) {
        ret = so_solve(p, tmp, xtol, maxfev, &nfev, 0);
        calls[0] += 1; calls[1] += nfev;
        if (ret != 1) {
            if (fabs(b - b_prec) < step_min) condition = 0;
            b = b_prec + (b - b_prec) / 2;
            memcpy(tmp, tab_param, sizeof(double) * P);
            *Rdata = (1 - b) * Rstart + b * goal;
        }
        if (ret == 1) {
            if (b == 1) condition = 0;
            else {
                b_prec = b;
                b = (b + bStep < 1.0) ? b + bStep : 1.0;
                memcpy(tab_param, tmp, sizeof(double) * P);
                *Rdata = (1 - b) * Rstart + b * goal;
            }
        }
    }
    if (ret == 1) memcpy(tab_param, tmp, sizeof(double) * P);
    free(tmp);
    return ret;
}

static void blend_boundary(so_problem *p, double b, const double *time_prec,
                           const double (*X_prec)[SO_MAX_DIM], const double *timed,
                           const double (*Xd)[SO_MAX_DIM])
{
    for (int i = 0; i <= p->num_multi; ++i) {
        p->time[i] = (1 - b) * time_prec[i] + b * timed[i];
        for (int j = 0; j < p->dim; ++j) p->Xb[i][j] = (1 - b) * X_prec[i][j] + b * Xd[i][j];
    }
}

/* shooting.cpp:598-692 */
int so_continuation_boundary(so_problem *p, double *tab_param, double xtol, int maxfev, double step,
                             const double *time_prec, const double (*X_prec)[SO_MAX_DIM],
                             const double *timed, const double (*Xd)[SO_MAX_DIM],
                             double step_min, int *calls)
{
    int P = so_num_param(p);
    double bStep = step;
    double b = step < 1.0 ? step : 1.0;
    double b_prec = 0;
    blend_boundary(p, b, time_prec, X_prec, timed, Xd);
    double *tmp = (double *)malloc(sizeof(double) * P);
    memcpy(tmp, tab_param, sizeof(double) * P);
    int condition = 1, ret = 0, nfev = 0;
    calls[0] = calls[1] = 0;
    while (condition) {
        ret = so_solve(p, tmp, xtol, maxfev, &nfev, 0);
        calls[0] += 1; calls[1] += nfev;
        if (ret != 1) {
            if (fabs(b - b_prec) < step_min) condition = 0;
            b = b_prec + (b - b_prec) / 2;
            memcpy(tmp, tab_param, sizeof(double) * P);
            blend_boundary(p, b, time_prec, X_prec, timed, Xd);
        }
        if (ret == 1) {
            if (b == 1) condition = 0;
            else {
                b_prec = b;
                b = (b + bStep < 1.0) ? b + bStep : 1.0;
                memcpy(tab_param, tmp, sizeof(double) * P);
                blend_boundary(p, b, time_prec, X_prec, timed, Xd);
            }
        }
    }
    if (ret == 1) memcpy(tab_param, tmp, sizeof(double) * P);
    free(tmp);
    return ret;
}
