/*
 * oracle/minpack.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Clean-room C restatement of the MINPACK Powell-hybrid solvers `hybrd` and
 * `hybrj` (More, Garbow, Hillstrom, "User Guide for MINPACK-1", ANL-80-74,
 * and the netlib Fortran it documents), exposed with the C signatures that the
 * reference binds at /root/reference/src/socp/shooting.cpp:803-826 (hybrd) and
 * :830-851 (hybrj).  The reference takes these two symbols from the third-party
 * library cminpack (github.com/devernay/cminpack, fetched UNPINNED by
 * /root/reference/src/socp/CMakeLists.txt:11-24, not vendored, not installed
 * in this image).  Parity pin: cross-checked against scipy.optimize._minpack
 * (same MINPACK lineage) in tests/test_oracle_minpack.py.
 */
#ifndef SOCP_ORACLE_MINPACK_H
#define SOCP_ORACLE_MINPACK_H

#ifdef __cplusplus
extern "C" {
#endif

/* callback ABI of cminpack, as used by shooting.hpp:282 and :293 */
typedef int (*minpack_func_nn)(void *p, int n, const double *x, double *fvec, int iflag);
typedef int (*minpack_funcder_nn)(void *p, int n, const double *x, double *fvec,
                                  double *fjac, int ldfjac, int iflag);

int hybrd(minpack_func_nn fcn, void *p, int n, double *x, double *fvec, double xtol,
          int maxfev, int ml, int mu, double epsfcn, double *diag, int mode,
          double factor, int nprint, int *nfev, double *fjac, int ldfjac, double *r,
          int lr, double *qtf, double *wa1, double *wa2, double *wa3, double *wa4);

int hybrj(minpack_funcder_nn fcn, void *p, int n, double *x, double *fvec, double *fjac,
          int ldfjac, double xtol, int maxfev, double *diag, int mode, double factor,
          int nprint, int *nfev, int *njev, double *r, int lr, double *qtf,
          double *wa1, double *wa2, double *wa3, double *wa4);

/* building blocks, exported so tests can pin them one by one */
double mp_enorm(int n, const double *x);
void mp_qrfac(int m, int n, double *a, int lda, double *rdiag, double *acnorm, double *wa);
void mp_qform(int m, int n, double *q, int ldq, double *wa);
void mp_dogleg(int n, const double *r, int lr, const double *diag, const double *qtb,
               double delta, double *x, double *wa1, double *wa2);
void mp_r1updt(int m, int n, double *s, int ls, const double *u, double *v, double *w, int *sing);
void mp_r1mpyq(int m, int n, double *a, int lda, const double *v, const double *w);

#ifdef __cplusplus
}
#endif
#endif
