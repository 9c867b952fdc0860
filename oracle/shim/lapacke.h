/*
 * oracle/shim/lapacke.h -- TEST INFRASTRUCTURE ONLY (never part of the product path).
 *
 * Prototype-only stand-in for <lapacke.h>, which this image does not ship.
 * It lets the reference's 2D translation units compile from where they lie
 * under /root/reference (see oracle/Makefile).  The symbols are provided at
 * link time either by oracle/mini_lapack.c (our restatement of the handful of
 * LAPACK routines the hot path calls, n <= 5) or by a real LAPACK (the OpenBLAS
 * bundled in the opencv wheel) for cross-checking.
 *
 * Third-party dependency being stood in for: LAPACK / LAPACKE, version
 * unpinned by the reference (nl-partsol/CMakeLists.txt:11,15,70-78).
 */
#ifndef NLPS_ORACLE_LAPACKE_SHIM_H
#define NLPS_ORACLE_LAPACKE_SHIM_H

#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102

typedef int lapack_int;

#ifdef __cplusplus
extern "C" {
#endif

lapack_int LAPACKE_dsyev(int layout, char jobz, char uplo, lapack_int n,
                         double *a, lapack_int lda, double *w);
lapack_int LAPACKE_dgetrf(int layout, lapack_int m, lapack_int n, double *a,
                          lapack_int lda, lapack_int *ipiv);
lapack_int LAPACKE_dgetrs(int layout, char trans, lapack_int n,
                          lapack_int nrhs, const double *a, lapack_int lda,
                          const lapack_int *ipiv, double *b, lapack_int ldb);
lapack_int LAPACKE_dgetri(int layout, lapack_int n, double *a, lapack_int lda,
                          const lapack_int *ipiv);
lapack_int LAPACKE_dgesv(int layout, lapack_int n, lapack_int nrhs, double *a,
                         lapack_int lda, lapack_int *ipiv, double *b,
                         lapack_int ldb);
double LAPACKE_dlange(int layout, char norm, lapack_int m, lapack_int n,
                      const double *a, lapack_int lda);
lapack_int LAPACKE_dgecon(int layout, char norm, lapack_int n, const double *a,
                          lapack_int lda, double anorm, double *rcond);

/* Fortran entry points the reference declares itself in some TUs and expects
 * from the LAPACK library in others. */
void dgetrf_(int *m, int *n, double *a, int *lda, int *ipiv, int *info);
void dgetri_(int *n, double *a, int *lda, int *ipiv, double *work, int *lwork,
             int *info);

#ifdef __cplusplus
}
#endif
#endif
