/*
 * oracle/minpack.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Clean-room restatement of MINPACK-1 `hybrd` / `hybrj` (Powell hybrid method)
 * and the routines they use (enorm, fdjac1, qrfac, qform, dogleg, r1updt,
 * r1mpyq), written from the published algorithm (ANL-80-74 and the netlib
 * Fortran documentation).  See minpack.h for provenance and the reference call
 * sites this replaces (/root/reference/src/socp/shooting.cpp:803-851).
 *
 * All matrices are column-major with leading dimension given, exactly as
 * MINPACK: a(i,j) == a[i + j*lda], 0-based here.
 * The packed upper-triangular factor `r` is stored BY ROWS (row 0 first).
 */
#include <math.h>
#include <float.h>
#include "minpack.h"

#define MP_EPSMCH DBL_EPSILON /* dpmpar(1) */
#define MP_DWARF  DBL_MIN     /* dpmpar(2) */
#define MP_GIANT  DBL_MAX     /* dpmpar(3) */

static double dmax(double a, double b) { return a >= b ? a : b; }
static double dmin(double a, double b) { return a <= b ? a : b; }

/* Euclidean norm with the three-accumulator over/underflow guard. */
double mp_enorm(int n, const double *x)
{
    const double rdwarf = 3.834e-20, rgiant = 1.304e19;
    double s1 = 0., s2 = 0., s3 = 0., x1max = 0., x3max = 0.;
    const double agiant = rgiant / (double)n;
    for (int i = 0; i < n; ++i) {
        double xabs = fabs(x[i]);
        if (xabs > rdwarf && xabs < agiant) {
            s2 += xabs * xabs;                      /* intermediate components */
        } else if (xabs <= rdwarf) {                /* small components */
            if (xabs > x3max) {
                double q = x3max / xabs;
                s3 = 1. + s3 * (q * q);
                x3max = xabs;
            } else if (xabs != 0.) {
                double q = xabs / x3max;
                s3 += q * q;
            }
        } else {                                    /* large components */
            if (xabs > x1max) {
                double q = x1max / xabs;
                s1 = 1. + s1 * (q * q);
                x1max = xabs;
            } else {
                double q = xabs / x1max;
                s1 += q * q;
            }
        }
    }
    if (s1 != 0.)
        return x1max * sqrt(s1 + (s2 / x1max) / x1max);
    if (s2 != 0.) {
        if (s2 >= x3max)
            return sqrt(s2 * (1. + (x3max / s2) * (x3max * s3)));
        return sqrt(x3max * ((s2 / x3max) + (x3max * s3)));
    }
    return x3max * sqrt(s3);
}

/* Forward-difference Jacobian; dense path when ml+mu+1 >= n, banded otherwise. */
static int mp_fdjac1(minpack_func_nn fcn, void *p, int n, double *x, const double *fvec,
                     double *fjac, int ldfjac, int ml, int mu, double epsfcn,
                     double *wa1, double *wa2)
{
    const double eps = sqrt(dmax(epsfcn, MP_EPSMCH));
    const int msum = ml + mu + 1;
    int iflag = 0;
    if (msum >= n) {
        for (int j = 0; j < n; ++j) {
            double temp = x[j];
            double h = eps * fabs(temp);
            if (h == 0.) h = eps;
            x[j] = temp + h;
            iflag = fcn(p, n, x, wa1, 2);
            if (iflag < 0) return iflag;
            x[j] = temp;
            for (int i = 0; i < n; ++i)
                fjac[i + j * ldfjac] = (wa1[i] - fvec[i]) / h;
        }
        return 0;
    }
    for (int k = 0; k < msum; ++k) {
        for (int j = k; j < n; j += msum) {
            wa2[j] = x[j];
            double h = eps * fabs(wa2[j]);
            if (h == 0.) h = eps;
            x[j] = wa2[j] + h;
        }
        iflag = fcn(p, n, x, wa1, 2);
        if (iflag < 0) return iflag;
        for (int j = k; j < n; j += msum) {
            x[j] = wa2[j];
            double h = eps * fabs(wa2[j]);
            if (h == 0.) h = eps;
            for (int i = 0; i < n; ++i) {
                fjac[i + j * ldfjac] = 0.;
                if (i >= j - mu && i <= j + ml)
                    fjac[i + j * ldfjac] = (wa1[i] - fvec[i]) / h;
            }
        }
    }
    return 0;
}

/* Householder QR without column pivoting (the only mode hybrd/hybrj use). */
void mp_qrfac(int m, int n, double *a, int lda, double *rdiag, double *acnorm, double *wa)
{
    for (int j = 0; j < n; ++j) {
        acnorm[j] = mp_enorm(m, &a[j * lda]);
        rdiag[j] = acnorm[j];
        wa[j] = rdiag[j];
    }
    const int minmn = m < n ? m : n;
    for (int j = 0; j < minmn; ++j) {
        double ajnorm = mp_enorm(m - j, &a[j + j * lda]);
        if (ajnorm != 0.) {
            if (a[j + j * lda] < 0.) ajnorm = -ajnorm;
            for (int i = j; i < m; ++i) a[i + j * lda] /= ajnorm;
            a[j + j * lda] += 1.;
            for (int k = j + 1; k < n; ++k) {
                double sum = 0.;
                for (int i = j; i < m; ++i) sum += a[i + j * lda] * a[i + k * lda];
                double temp = sum / a[j + j * lda];
                for (int i = j; i < m; ++i) a[i + k * lda] -= temp * a[i + j * lda];
            }
        }
        rdiag[j] = -ajnorm;
    }
}

/* Accumulate the orthogonal factor from the Householder vectors left by qrfac. */
void mp_qform(int m, int n, double *q, int ldq, double *wa)
{
    const int minmn = m < n ? m : n;
    for (int j = 1; j < minmn; ++j)
        for (int i = 0; i < j; ++i) q[i + j * ldq] = 0.;
    for (int j = n; j < m; ++j) {
        for (int i = 0; i < m; ++i) q[i + j * ldq] = 0.;
        q[j + j * ldq] = 1.;
    }
    for (int l = 0; l < minmn; ++l) {
        int k = minmn - 1 - l;
        for (int i = k; i < m; ++i) {
            wa[i] = q[i + k * ldq];
            q[i + k * ldq] = 0.;
        }
        q[k + k * ldq] = 1.;
        if (wa[k] != 0.) {
            for (int j = k; j < m; ++j) {
                double sum = 0.;
                for (int i = k; i < m; ++i) sum += q[i + j * ldq] * wa[i];
                double temp = sum / wa[k];
                for (int i = k; i < m; ++i) q[i + j * ldq] -= temp * wa[i];
            }
        }
    }
}

/* Dogleg step: combination of Gauss-Newton and scaled-gradient directions. */
void mp_dogleg(int n, const double *r, int lr, const double *diag, const double *qtb,
               double delta, double *x, double *wa1, double *wa2)
{
    (void)lr;
    /* Gauss-Newton direction by back substitution on the packed R */
    int jj = (n * (n + 1)) / 2;             /* one past the last packed element */
    for (int k = 1; k <= n; ++k) {
        int j = n - k;
        jj -= k;                            /* index of r(j,j) */
        int l = jj + 1;
        double sum = 0.;
        for (int i = j + 1; i < n; ++i) { sum += r[l] * x[i]; ++l; }
        double temp = r[jj];
        if (temp == 0.) {
            l = j;
            for (int i = 0; i <= j; ++i) {
                temp = dmax(temp, fabs(r[l]));
                l += n - 1 - i;
            }
            temp = MP_EPSMCH * temp;
            if (temp == 0.) temp = MP_EPSMCH;
        }
        x[j] = (qtb[j] - sum) / temp;
    }
    for (int j = 0; j < n; ++j) { wa1[j] = 0.; wa2[j] = diag[j] * x[j]; }
    double qnorm = mp_enorm(n, wa2);
    if (qnorm <= delta) return;

    /* scaled gradient direction */
    int l = 0;
    for (int j = 0; j < n; ++j) {
        double temp = qtb[j];
        for (int i = j; i < n; ++i) { wa1[i] += r[l] * temp; ++l; }
        wa1[j] /= diag[j];
    }
    double gnorm = mp_enorm(n, wa1);
    double sgnorm = 0.;
    double alpha = delta / qnorm;
    if (gnorm != 0.) {
        for (int j = 0; j < n; ++j) wa1[j] = (wa1[j] / gnorm) / diag[j];
        l = 0;
        for (int j = 0; j < n; ++j) {
            double sum = 0.;
            for (int i = j; i < n; ++i) { sum += r[l] * wa1[i]; ++l; }
            wa2[j] = sum;
        }
        double temp = mp_enorm(n, wa2);
        sgnorm = (gnorm / temp) / temp;
        alpha = 0.;
        if (sgnorm < delta) {
            double bnorm = mp_enorm(n, qtb);
            temp = (bnorm / gnorm) * (bnorm / qnorm) * (sgnorm / delta);
            double dq = delta / qnorm, sd = sgnorm / delta;
            temp = temp - dq * (sd * sd)
                 + sqrt((temp - dq) * (temp - dq) + (1. - dq * dq) * (1. - sd * sd));
            alpha = (dq * (1. - sd * sd)) / temp;
        }
    }
    double temp = (1. - alpha) * dmin(sgnorm, delta);
    for (int j = 0; j < n; ++j) x[j] = temp * wa1[j] + alpha * x[j];
}

/* Apply the 2(n-1) Givens rotations recorded in v, w (by r1updt) to A (m x n). */
void mp_r1mpyq(int m, int n, double *a, int lda, const double *v, const double *w)
{
    const int nm1 = n - 1;
    if (nm1 < 1) return;
    double c, s;
    for (int nmj = 1; nmj <= nm1; ++nmj) {
        int j = n - 1 - nmj;
        if (fabs(v[j]) > 1.) { c = 1. / v[j]; s = sqrt(1. - c * c); }
        else                 { s = v[j];      c = sqrt(1. - s * s); }
        for (int i = 0; i < m; ++i) {
            double temp = c * a[i + j * lda] - s * a[i + nm1 * lda];
            a[i + nm1 * lda] = s * a[i + j * lda] + c * a[i + nm1 * lda];
            a[i + j * lda] = temp;
        }
    }
    for (int j = 0; j < nm1; ++j) {
        if (fabs(w[j]) > 1.) { c = 1. / w[j]; s = sqrt(1. - c * c); }
        else                 { s = w[j];      c = sqrt(1. - s * s); }
        for (int i = 0; i < m; ++i) {
            double temp = c * a[i + j * lda] + s * a[i + nm1 * lda];
            a[i + nm1 * lda] = -s * a[i + j * lda] + c * a[i + nm1 * lda];
            a[i + j * lda] = temp;
        }
    }
}

/* Rank-1 update of the packed LOWER-trapezoidal s (m x n, stored by columns):
 * finds orthogonal Q with (s + u v^T) Q lower trapezoidal again. */
void mp_r1updt(int m, int n, double *s, int ls, const double *u, double *v, double *w, int *sing)
{
    (void)ls;
    const double p5 = .5, p25 = .25;
    const double giant = MP_GIANT;
    /* 1-based bookkeeping as in the published algorithm, shifted at the accesses */
    int jj = (n * (2 * m - n + 1)) / 2 - (m - n);   /* 1-based index of s(n,n) */
    int l = jj;
    for (int i = n; i <= m; ++i) { w[i - 1] = s[l - 1]; ++l; }

    const int nm1 = n - 1;
    for (int nmj = 1; nmj <= nm1; ++nmj) {
        int j = n - nmj;
        jj -= (m - j + 1);
        w[j - 1] = 0.;
        if (v[j - 1] != 0.) {
            double c, sn, tau;
            if (fabs(v[n - 1]) < fabs(v[j - 1])) {
                double cotan = v[n - 1] / v[j - 1];
                sn = p5 / sqrt(p25 + p25 * (cotan * cotan));
                c = sn * cotan;
                tau = 1.;
                if (fabs(c) * giant > 1.) tau = 1. / c;
            } else {
                double tn = v[j - 1] / v[n - 1];
                c = p5 / sqrt(p25 + p25 * (tn * tn));
                sn = c * tn;
                tau = sn;
            }
            v[n - 1] = sn * v[j - 1] + c * v[n - 1];
            v[j - 1] = tau;
            l = jj;
            for (int i = j; i <= m; ++i) {
                double temp = c * s[l - 1] - sn * w[i - 1];
                w[i - 1] = sn * s[l - 1] + c * w[i - 1];
                s[l - 1] = temp;
                ++l;
            }
        }
    }
    for (int i = 1; i <= m; ++i) w[i - 1] += v[n - 1] * u[i - 1];

    *sing = 0;
    for (int j = 1; j <= nm1; ++j) {
        if (w[j - 1] != 0.) {
            double c, sn, tau;
            if (fabs(s[jj - 1]) < fabs(w[j - 1])) {
                double cotan = s[jj - 1] / w[j - 1];
                sn = p5 / sqrt(p25 + p25 * (cotan * cotan));
                c = sn * cotan;
                tau = 1.;
                if (fabs(c) * giant > 1.) tau = 1. / c;
            } else {
                double tn = w[j - 1] / s[jj - 1];
                c = p5 / sqrt(p25 + p25 * (tn * tn));
                sn = c * tn;
                tau = sn;
            }
            l = jj;
            for (int i = j; i <= m; ++i) {
                double temp = c * s[l - 1] + sn * w[i - 1];
                w[i - 1] = -sn * s[l - 1] + c * w[i - 1];
                s[l - 1] = temp;
                ++l;
            }
            w[j - 1] = tau;
        }
        if (s[jj - 1] == 0.) *sing = 1;
        jj += (m - j + 1);
    }
    l = jj;
    for (int i = n; i <= m; ++i) { s[l - 1] = w[i - 1]; ++l; }
    if (s[jj - 1] == 0.) *sing = 1;
}

/* ---- shared body of hybrd / hybrj ------------------------------------------------ */

typedef struct {
    minpack_func_nn f;        /* hybrd callback, or NULL */
    minpack_funcder_nn fj;    /* hybrj callback, or NULL */
    void *p;
} mp_cb;

static int mp_eval(const mp_cb *cb, int n, const double *x, double *fvec, double *fjac, int ldfjac)
{
    if (cb->f) return cb->f(cb->p, n, x, fvec, 1);
    return cb->fj(cb->p, n, x, fvec, fjac, ldfjac, 1);
}

static int mp_hybrid(const mp_cb *cb, int n, double *x, double *fvec, double xtol, int maxfev,
                     int ml, int mu, double epsfcn, double *diag, int mode, double factor,
                     int *nfev, int *njev, double *fjac, int ldfjac, double *r, int lr,
                     double *qtf, double *wa1, double *wa2, double *wa3, double *wa4)
{
    const double p1 = .1, p5 = .5, p001 = .001, p0001 = 1e-4;
    const double epsmch = MP_EPSMCH;
    int info = 0, iflag = 0;
    *nfev = 0;
    if (njev) *njev = 0;

    if (n <= 0 || xtol < 0. || maxfev <= 0 || factor <= 0. || ldfjac < n || lr < (n * (n + 1)) / 2)
        return 0;
    if (cb->f && (ml < 0 || mu < 0)) return 0;
    if (mode == 2)
        for (int j = 0; j < n; ++j)
            if (diag[j] <= 0.) return 0;

    iflag = mp_eval(cb, n, x, fvec, fjac, ldfjac);
    *nfev = 1;
    if (iflag < 0) return iflag;
    double fnorm = mp_enorm(n, fvec);

    int msum = ml + mu + 1;
    if (msum > n) msum = n;

    int iter = 1, ncsuc = 0, ncfail = 0, nslow1 = 0, nslow2 = 0;
    double delta = 0., xnorm = 0.;

    for (;;) {                                          /* outer loop: fresh Jacobian */
        int jeval = 1;
        if (cb->f) {
            iflag = mp_fdjac1(cb->f, cb->p, n, x, fvec, fjac, ldfjac, ml, mu, epsfcn, wa1, wa2);
            *nfev += msum;
        } else {
            iflag = cb->fj(cb->p, n, x, fvec, fjac, ldfjac, 2);
            ++(*njev);
        }
        if (iflag < 0) return iflag;

        mp_qrfac(n, n, fjac, ldfjac, wa1, wa2, wa3);

        if (iter == 1) {
            if (mode != 2)
                for (int j = 0; j < n; ++j) {
                    diag[j] = wa2[j];
                    if (wa2[j] == 0.) diag[j] = 1.;
                }
            for (int j = 0; j < n; ++j) wa3[j] = diag[j] * x[j];
            xnorm = mp_enorm(n, wa3);
            delta = factor * xnorm;
            if (delta == 0.) delta = factor;
        }

        /* qtf = Q^T fvec from the Householder vectors */
        for (int i = 0; i < n; ++i) qtf[i] = fvec[i];
        for (int j = 0; j < n; ++j) {
            if (fjac[j + j * ldfjac] != 0.) {
                double sum = 0.;
                for (int i = j; i < n; ++i) sum += fjac[i + j * ldfjac] * qtf[i];
                double temp = -sum / fjac[j + j * ldfjac];
                for (int i = j; i < n; ++i) qtf[i] += fjac[i + j * ldfjac] * temp;
            }
        }

        /* copy R (upper triangle, by rows) into the packed array */
        int sing = 0;
        for (int j = 0; j < n; ++j) {
            int l = j;
            for (int i = 0; i < j; ++i) {
                r[l] = fjac[i + j * ldfjac];
                l += n - 1 - i;
            }
            r[l] = wa1[j];
            if (wa1[j] == 0.) sing = 1;
        }
        (void)sing;

        mp_qform(n, n, fjac, ldfjac, wa1);

        if (mode != 2)
            for (int j = 0; j < n; ++j) diag[j] = dmax(diag[j], wa2[j]);

        for (;;) {                                      /* inner loop: Broyden updates */
            mp_dogleg(n, r, lr, diag, qtf, delta, wa1, wa2, wa3);

            for (int j = 0; j < n; ++j) {
                wa1[j] = -wa1[j];
                wa2[j] = x[j] + wa1[j];
                wa3[j] = diag[j] * wa1[j];
            }
            double pnorm = mp_enorm(n, wa3);
            if (iter == 1) delta = dmin(delta, pnorm);

            iflag = mp_eval(cb, n, wa2, wa4, fjac, ldfjac);
            ++(*nfev);
            if (iflag < 0) return iflag;
            double fnorm1 = mp_enorm(n, wa4);

            double actred = -1.;
            if (fnorm1 < fnorm) { double q = fnorm1 / fnorm; actred = 1. - q * q; }

            int l = 0;
            for (int i = 0; i < n; ++i) {
                double sum = 0.;
                for (int j = i; j < n; ++j) { sum += r[l] * wa1[j]; ++l; }
                wa3[i] = qtf[i] + sum;
            }
            double temp = mp_enorm(n, wa3);
            double prered = 0.;
            if (temp < fnorm) { double q = temp / fnorm; prered = 1. - q * q; }

            double ratio = 0.;
            if (prered > 0.) ratio = actred / prered;

            if (ratio < p1) {
                ncsuc = 0;
                ++ncfail;
                delta = p5 * delta;
            } else {
                ncfail = 0;
                ++ncsuc;
                if (ratio >= p5 || ncsuc > 1) delta = dmax(delta, pnorm / p5);
                if (fabs(ratio - 1.) <= p1) delta = pnorm / p5;
            }

            if (ratio >= p0001) {
                for (int j = 0; j < n; ++j) {
                    x[j] = wa2[j];
                    wa2[j] = diag[j] * x[j];
                    fvec[j] = wa4[j];
                }
                xnorm = mp_enorm(n, wa2);
                fnorm = fnorm1;
                ++iter;
            }

            ++nslow1;
            if (actred >= p001) nslow1 = 0;
            if (jeval) ++nslow2;
            if (actred >= p1) nslow2 = 0;

            if (delta <= xtol * xnorm || fnorm == 0.) info = 1;
            if (info != 0) return info;

            if (*nfev >= maxfev) info = 2;
            if (p1 * dmax(p1 * delta, pnorm) <= epsmch * xnorm) info = 3;
            if (nslow2 == 5) info = 4;
            if (nslow1 == 10) info = 5;
            if (info != 0) return info;

            if (ncfail == 2) break;                     /* re-evaluate the Jacobian */

            /* rank-one (Broyden) modification of the QR factors */
            for (int j = 0; j < n; ++j) {
                double sum = 0.;
                for (int i = 0; i < n; ++i) sum += fjac[i + j * ldfjac] * wa4[i];
                wa2[j] = (sum - wa3[j]) / pnorm;
                wa1[j] = diag[j] * ((diag[j] * wa1[j]) / pnorm);
                if (ratio >= p0001) qtf[j] = sum;
            }
            mp_r1updt(n, n, r, lr, wa1, wa2, wa3, &sing);
            mp_r1mpyq(n, n, fjac, ldfjac, wa2, wa3);
            mp_r1mpyq(1, n, qtf, 1, wa2, wa3);
            jeval = 0;
        }
    }
}

int hybrd(minpack_func_nn fcn, void *p, int n, double *x, double *fvec, double xtol,
          int maxfev, int ml, int mu, double epsfcn, double *diag, int mode,
          double factor, int nprint, int *nfev, double *fjac, int ldfjac, double *r,
          int lr, double *qtf, double *wa1, double *wa2, double *wa3, double *wa4)
{
    (void)nprint;   /* the reference always passes nprint = 0 (shooting.cpp:101) */
    mp_cb cb = { fcn, 0, p };
    return mp_hybrid(&cb, n, x, fvec, xtol, maxfev, ml, mu, epsfcn, diag, mode, factor,
                     nfev, 0, fjac, ldfjac, r, lr, qtf, wa1, wa2, wa3, wa4);
}

int hybrj(minpack_funcder_nn fcn, void *p, int n, double *x, double *fvec, double *fjac,
          int ldfjac, double xtol, int maxfev, double *diag, int mode, double factor,
          int nprint, int *nfev, int *njev, double *r, int lr, double *qtf,
          double *wa1, double *wa2, double *wa3, double *wa4)
{
    (void)nprint;
    mp_cb cb = { 0, fcn, p };
    return mp_hybrid(&cb, n, x, fvec, xtol, maxfev, 0, 0, 0., diag, mode, factor,
                     nfev, njev, fjac, ldfjac, r, lr, qtf, wa1, wa2, wa3, wa4);
}
