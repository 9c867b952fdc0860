/*
 * oracle/mini_lapack.c -- TEST INFRASTRUCTURE ONLY (never part of the product path).
 *
 * Restatement of the few LAPACK / LAPACKE routines NL-PartSol's hot path calls
 * (n <= 5), so that the reference's own 2D sources can be linked into
 * oracle/_ref without an external LAPACK, and so that the C oracle port and the
 * CUDA kernels have a pinned definition of the third-party arithmetic.
 *
 * Dependency restated: LAPACK/LAPACKE (version unpinned by the reference,
 * nl-partsol/CMakeLists.txt:11,15,70-78; the author's machines used OpenBLAS).
 * Algorithms follow the published reference-LAPACK routines:
 *   DSYEV  -> DSYTRD + DORGTR + DSTEQR; for n == 2 the tridiagonalisation is
 *             the identity and DSTEQR reduces to one DLAEV2 rotation followed
 *             by its ascending selection sort.  That path is restated exactly
 *             (eigenvector SIGNS included) because NL-PartSol's plastic
 *             branches index the eigenvector matrix transposed
 *             (Constitutive/Plasticity/Drucker-Prager.c:957-958 vs :770-771),
 *             which makes the result depend on LAPACK's sign convention.
 *   DGETRF -> partial-pivoting LU (first maximal |a_ik| wins, as IDAMAX).
 *   DGETRS / DGETRI -> triangular solves with the factors.
 *   DLANGE('1'), DGECON('1') -> 1-norm and reciprocal condition estimate.  The
 *             Hager/Higham estimator of DGECON is replaced by the exact 1-norm
 *             of inv(L*U) (n <= 5), documented deviation: the reference only
 *             compares the result with 1e-8 (Nodes/LME.c:308).
 * Call sites in the reference: Drucker-Prager.c:635, Matsuoka-Nakai.c:723,
 * 1100-1154, 1242-1264, TensorLib.c:208,783-886,981-984, MatrixOp.c:341-359,
 * compute-Strains.c:286-308.
 *
 * tests/test_mini_lapack.py cross-checks every routine against the OpenBLAS
 * 0.3.15 LAPACK bundled in the opencv wheel of this image when it is present.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "shim/lapacke.h"

#define MINI_NMAX 8

/* ---------------------------------------------------------------------- */
/* DLAEV2: eigen-decomposition of [[a,b],[b,c]] (reference LAPACK dlaev2.f) */
void nlps_dlaev2(double a, double b, double c, double *rt1, double *rt2,
                 double *cs1, double *sn1) {
  double sm = a + c, df = a - c, adf = fabs(df), tb = b + b, ab = fabs(tb);
  double acmx, acmn, rt, cs, ct, tn, acs;
  int sgn1, sgn2;
  if (fabs(a) > fabs(c)) {
    acmx = a;
    acmn = c;
  } else {
    acmx = c;
    acmn = a;
  }
  if (adf > ab) {
    double q = ab / adf;
    rt = adf * sqrt(1.0 + q * q);
  } else if (adf < ab) {
    double q = adf / ab;
    rt = ab * sqrt(1.0 + q * q);
  } else {
    rt = ab * sqrt(2.0);
  }
  if (sm < 0.0) {
    *rt1 = 0.5 * (sm - rt);
    sgn1 = -1;
    *rt2 = (acmx / *rt1) * acmn - (b / *rt1) * b;
  } else if (sm > 0.0) {
    *rt1 = 0.5 * (sm + rt);
    sgn1 = 1;
    *rt2 = (acmx / *rt1) * acmn - (b / *rt1) * b;
  } else {
    *rt1 = 0.5 * rt;
    *rt2 = -0.5 * rt;
    sgn1 = 1;
  }
  if (df >= 0.0) {
    cs = df + rt;
    sgn2 = 1;
  } else {
    cs = df - rt;
    sgn2 = -1;
  }
  acs = fabs(cs);
  if (acs > ab) {
    ct = -tb / cs;
    *sn1 = 1.0 / sqrt(1.0 + ct * ct);
    *cs1 = ct * *sn1;
  } else {
    if (ab == 0.0) {
      *cs1 = 1.0;
      *sn1 = 0.0;
    } else {
      tn = -cs / tb;
      *cs1 = 1.0 / sqrt(1.0 + tn * tn);
      *sn1 = tn * *cs1;
    }
  }
  if (sgn1 == sgn2) {
    tn = *cs1;
    *cs1 = -*sn1;
    *sn1 = tn;
  }
}

/*
 * DSYEV('V') for a 2x2 symmetric matrix [[d1,e],[e,d2]], as DSYTRD (identity
 * for n=2) + DSTEQR would do it.  w ascending; z row-major, eigenvector j in
 * column j (what LAPACKE_dsyev(LAPACK_ROW_MAJOR,...) hands back).
 */
void nlps_dsyev2(double d1, double e, double d2, double w[2], double z[4]) {
  const double eps = 0x1p-53;                 /* dlamch('E') */
  const double eps2 = eps * eps;
  const double safmin = DBL_MIN;              /* dlamch('S') */
  double z11 = 1.0, z12 = 0.0, z21 = 0.0, z22 = 1.0;
  double tst = fabs(e);
  int split = 0;
  /* dsteqr.f: initial scan for a negligible off-diagonal */
  if (tst == 0.0) {
    split = 1;
  } else if (tst <= (sqrt(fabs(d1)) * sqrt(fabs(d2))) * eps) {
    split = 1;
  }
  if (!split) {
    /* QL/QR convergence test on the (only) off-diagonal */
    double t2 = fabs(e) * fabs(e);
    if (t2 <= (eps2 * fabs(d1)) * fabs(d2) + safmin) split = 1;
  }
  if (!split) {
    double rt1, rt2, c, s;
    nlps_dlaev2(d1, e, d2, &rt1, &rt2, &c, &s);
    /* dlasr('R','V',.,n,2,c,s,Z) applied to Z = I */
    z11 = c;
    z12 = -s;
    z21 = s;
    z22 = c;
    d1 = rt1;
    d2 = rt2;
  }
  /* ascending selection sort with column swap */
  if (d2 < d1) {
    double t = d1;
    d1 = d2;
    d2 = t;
    t = z11; z11 = z12; z12 = t;
    t = z21; z21 = z22; z22 = t;
  }
  w[0] = d1;
  w[1] = d2;
  z[0] = z11;
  z[1] = z12;
  z[2] = z21;
  z[3] = z22;
}

/* Cyclic Jacobi for 3 <= n <= MINI_NMAX (row-major a, full symmetric).  Not
 * sign-faithful to DSTEQR; used only where the reference has no compilable
 * 3D path to be faithful to.  Eigenvalues ascending, eigenvector j in
 * column j, each eigenvector normalised so that its largest-|.| component
 * is positive (our pinned convention). */
void nlps_jacobi_eig(int n, const double *a_in, double *w, double *z) {
  double a[MINI_NMAX * MINI_NMAX];
  int i, j, k, sweep;
  for (i = 0; i < n; i++)
    for (j = 0; j < n; j++) {
      a[i * n + j] = (j >= i) ? a_in[i * n + j] : a_in[j * n + i];
      z[i * n + j] = (i == j) ? 1.0 : 0.0;
    }
  for (sweep = 0; sweep < 64; sweep++) {
    double off = 0.0, diag = 0.0;
    for (i = 0; i < n; i++) {
      diag += a[i * n + i] * a[i * n + i];
      for (j = i + 1; j < n; j++) off += a[i * n + j] * a[i * n + j];
    }
    if (off <= 1e-34 * diag || off == 0.0) break;
    for (i = 0; i < n - 1; i++)
      for (j = i + 1; j < n; j++) {
        double apq = a[i * n + j];
        if (apq == 0.0) continue;
        double theta = (a[j * n + j] - a[i * n + i]) / (2.0 * apq);
        double t = (theta >= 0.0 ? 1.0 : -1.0) /
                   (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (k = 0; k < n; k++) {
          double akp = a[k * n + i], akq = a[k * n + j];
          a[k * n + i] = c * akp - s * akq;
          a[k * n + j] = s * akp + c * akq;
        }
        for (k = 0; k < n; k++) {
          double apk = a[i * n + k], aqk = a[j * n + k];
          a[i * n + k] = c * apk - s * aqk;
          a[j * n + k] = s * apk + c * aqk;
        }
        for (k = 0; k < n; k++) {
          double zkp = z[k * n + i], zkq = z[k * n + j];
          z[k * n + i] = c * zkp - s * zkq;
          z[k * n + j] = s * zkp + c * zkq;
        }
      }
  }
  for (i = 0; i < n; i++) w[i] = a[i * n + i];
  for (i = 0; i < n - 1; i++) {
    k = i;
    for (j = i + 1; j < n; j++)
      if (w[j] < w[k]) k = j;
    if (k != i) {
      double t = w[i];
      w[i] = w[k];
      w[k] = t;
      for (j = 0; j < n; j++) {
        t = z[j * n + i];
        z[j * n + i] = z[j * n + k];
        z[j * n + k] = t;
      }
    }
  }
  for (j = 0; j < n; j++) {
    int kmax = 0;
    for (i = 1; i < n; i++)
      if (fabs(z[i * n + j]) > fabs(z[kmax * n + j])) kmax = i;
    if (z[kmax * n + j] < 0.0)
      for (i = 0; i < n; i++) z[i * n + j] = -z[i * n + j];
  }
}

lapack_int LAPACKE_dsyev(int layout, char jobz, char uplo, lapack_int n,
                         double *a, lapack_int lda, double *w) {
  (void)jobz;
  if (n < 1 || n > MINI_NMAX) return -4;
  if (n == 1) {
    w[0] = a[0];
    a[0] = 1.0;
    return 0;
  }
  /* logical (i,j), i<=j for 'U', i>=j for 'L' */
  int upper = (uplo == 'U' || uplo == 'u');
#define AIJ(i, j) (layout == LAPACK_ROW_MAJOR ? a[(i) * lda + (j)] : a[(j) * lda + (i)])
  if (n == 2) {
    double z[4], ww[2];
    double e = upper ? AIJ(0, 1) : AIJ(1, 0);
    nlps_dsyev2(AIJ(0, 0), e, AIJ(1, 1), ww, z);
    w[0] = ww[0];
    w[1] = ww[1];
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++) {
        if (layout == LAPACK_ROW_MAJOR) a[i * lda + j] = z[i * 2 + j];
        else a[j * lda + i] = z[i * 2 + j];
      }
    return 0;
  }
  {
    double s[MINI_NMAX * MINI_NMAX], z[MINI_NMAX * MINI_NMAX];
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        int ii = i, jj = j;
        if ((upper && i > j) || (!upper && i < j)) { ii = j; jj = i; }
        s[i * n + j] = AIJ(ii, jj);
      }
    nlps_jacobi_eig(n, s, w, z);
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        if (layout == LAPACK_ROW_MAJOR) a[i * lda + j] = z[i * n + j];
        else a[j * lda + i] = z[i * n + j];
      }
  }
#undef AIJ
  return 0;
}

/* ---------------------------------------------------------------------- */
/* LU with partial pivoting on a column-major n x n matrix (dgetf2).       */
static int lu_colmajor(int m, int n, double *a, int lda, int *ipiv) {
  int info = 0, mn = m < n ? m : n;
  for (int j = 0; j < mn; j++) {
    int jp = j;
    double amax = fabs(a[j * lda + j]);
    for (int i = j + 1; i < m; i++)
      if (fabs(a[j * lda + i]) > amax) {
        amax = fabs(a[j * lda + i]);
        jp = i;
      }
    ipiv[j] = jp + 1;
    if (a[j * lda + jp] != 0.0) {
      if (jp != j)
        for (int k = 0; k < n; k++) {
          double t = a[k * lda + j];
          a[k * lda + j] = a[k * lda + jp];
          a[k * lda + jp] = t;
        }
      double piv = a[j * lda + j];
      if (fabs(piv) >= DBL_MIN) {
        double r = 1.0 / piv;
        for (int i = j + 1; i < m; i++) a[j * lda + i] *= r;
      } else {
        for (int i = j + 1; i < m; i++) a[j * lda + i] /= piv;
      }
    } else if (info == 0) {
      info = j + 1;
    }
    for (int k = j + 1; k < n; k++) {
      double akj = a[k * lda + j];
      for (int i = j + 1; i < m; i++) a[k * lda + i] -= a[j * lda + i] * akj;
    }
  }
  return info;
}

void dgetrf_(int *m, int *n, double *a, int *lda, int *ipiv, int *info) {
  *info = lu_colmajor(*m, *n, a, *lda, ipiv);
}

/* solve (P^T L U) x = b in place, column-major factors */
static void lu_solve_colmajor(int n, const double *a, int lda, const int *ipiv,
                              double *b) {
  for (int i = 0; i < n; i++) {
    int p = ipiv[i] - 1;
    if (p != i) {
      double t = b[i];
      b[i] = b[p];
      b[p] = t;
    }
  }
  for (int j = 0; j < n; j++)
    for (int i = j + 1; i < n; i++) b[i] -= a[j * lda + i] * b[j];
  for (int j = n - 1; j >= 0; j--) {
    b[j] /= a[j * lda + j];
    for (int i = 0; i < j; i++) b[i] -= a[j * lda + i] * b[j];
  }
}

/* solve A^T x = b with the factors of A: U^T L^T P x = b */
static void lu_solve_trans_colmajor(int n, const double *a, int lda,
                                    const int *ipiv, double *b) {
  for (int j = 0; j < n; j++) {
    for (int i = 0; i < j; i++) b[j] -= a[j * lda + i] * b[i];
    b[j] /= a[j * lda + j];
  }
  for (int j = n - 1; j >= 0; j--)
    for (int i = j + 1; i < n; i++) b[j] -= a[j * lda + i] * b[i];
  for (int i = n - 1; i >= 0; i--) {
    int p = ipiv[i] - 1;
    if (p != i) {
      double t = b[i];
      b[i] = b[p];
      b[p] = t;
    }
  }
}

void dgetri_(int *n, double *a, int *lda, int *ipiv, double *work, int *lwork,
             int *info) {
  (void)work;
  (void)lwork;
  int nn = *n, ld = *lda;
  double inv[MINI_NMAX * MINI_NMAX];
  *info = 0;
  for (int j = 0; j < nn; j++)
    if (a[j * ld + j] == 0.0) {
      *info = j + 1;
      return;
    }
  for (int j = 0; j < nn; j++) {
    double col[MINI_NMAX];
    for (int i = 0; i < nn; i++) col[i] = (i == j) ? 1.0 : 0.0;
    lu_solve_colmajor(nn, a, ld, ipiv, col);
    for (int i = 0; i < nn; i++) inv[j * nn + i] = col[i];
  }
  for (int j = 0; j < nn; j++)
    for (int i = 0; i < nn; i++) a[j * ld + i] = inv[j * nn + i];
}

lapack_int LAPACKE_dgetrf(int layout, lapack_int m, lapack_int n, double *a,
                          lapack_int lda, lapack_int *ipiv) {
  if (layout == LAPACK_COL_MAJOR) return lu_colmajor(m, n, a, lda, ipiv);
  double t[MINI_NMAX * MINI_NMAX];
  if (m > MINI_NMAX || n > MINI_NMAX) return -2;
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) t[j * m + i] = a[i * lda + j];
  int info = lu_colmajor(m, n, t, m, ipiv);
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) a[i * lda + j] = t[j * m + i];
  return info;
}

lapack_int LAPACKE_dgetrs(int layout, char trans, lapack_int n,
                          lapack_int nrhs, const double *a, lapack_int lda,
                          const lapack_int *ipiv, double *b, lapack_int ldb) {
  double t[MINI_NMAX * MINI_NMAX], col[MINI_NMAX];
  if (n > MINI_NMAX) return -3;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++)
      t[j * n + i] = (layout == LAPACK_ROW_MAJOR) ? a[i * lda + j] : a[j * lda + i];
  int tr = !(trans == 'N' || trans == 'n');
  for (int r = 0; r < nrhs; r++) {
    for (int i = 0; i < n; i++)
      col[i] = (layout == LAPACK_ROW_MAJOR) ? b[i * ldb + r] : b[r * ldb + i];
    if (tr) lu_solve_trans_colmajor(n, t, n, ipiv, col);
    else lu_solve_colmajor(n, t, n, ipiv, col);
    for (int i = 0; i < n; i++) {
      if (layout == LAPACK_ROW_MAJOR) b[i * ldb + r] = col[i];
      else b[r * ldb + i] = col[i];
    }
  }
  return 0;
}

lapack_int LAPACKE_dgetri(int layout, lapack_int n, double *a, lapack_int lda,
                          const lapack_int *ipiv) {
  double t[MINI_NMAX * MINI_NMAX], work[MINI_NMAX];
  int info, lw = MINI_NMAX, nn = n, ld = n;
  int piv[MINI_NMAX];
  for (int i = 0; i < n; i++) piv[i] = ipiv[i];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++)
      t[j * n + i] = (layout == LAPACK_ROW_MAJOR) ? a[i * lda + j] : a[j * lda + i];
  dgetri_(&nn, t, &ld, piv, work, &lw, &info);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      if (layout == LAPACK_ROW_MAJOR) a[i * lda + j] = t[j * n + i];
      else a[j * lda + i] = t[j * n + i];
    }
  return info;
}

lapack_int LAPACKE_dgesv(int layout, lapack_int n, lapack_int nrhs, double *a,
                         lapack_int lda, lapack_int *ipiv, double *b,
                         lapack_int ldb) {
  lapack_int info = LAPACKE_dgetrf(layout, n, n, a, lda, ipiv);
  if (info != 0) return info;
  return LAPACKE_dgetrs(layout, 'N', n, nrhs, a, lda, ipiv, b, ldb);
}

double LAPACKE_dlange(int layout, char norm, lapack_int m, lapack_int n,
                      const double *a, lapack_int lda) {
  double v = 0.0;
#define AIJ(i, j) (layout == LAPACK_ROW_MAJOR ? a[(i) * lda + (j)] : a[(j) * lda + (i)])
  if (norm == '1' || norm == 'O' || norm == 'o') {
    for (int j = 0; j < n; j++) {
      double s = 0.0;
      for (int i = 0; i < m; i++) s += fabs(AIJ(i, j));
      if (s > v) v = s;
    }
  } else if (norm == 'I' || norm == 'i') {
    for (int i = 0; i < m; i++) {
      double s = 0.0;
      for (int j = 0; j < n; j++) s += fabs(AIJ(i, j));
      if (s > v) v = s;
    }
  } else if (norm == 'M' || norm == 'm') {
    for (int i = 0; i < m; i++)
      for (int j = 0; j < n; j++)
        if (fabs(AIJ(i, j)) > v) v = fabs(AIJ(i, j));
  } else {
    for (int i = 0; i < m; i++)
      for (int j = 0; j < n; j++) v += AIJ(i, j) * AIJ(i, j);
    v = sqrt(v);
  }
#undef AIJ
  return v;
}

/*
 * DGECON interprets `a` as the L (unit lower) and U factors of a DGETRF call
 * and returns rcond = 1 / (anorm * ||inv(L*U)||).  The reference passes the
 * UNFACTORISED matrix (Matlib/TensorLib.c:981-984), so the number is the
 * condition estimate of the matrix whose LU factors happen to be the entries
 * of `a`; this routine reproduces that semantics with the exact norm.
 */
lapack_int LAPACKE_dgecon(int layout, char norm, lapack_int n, const double *a,
                          lapack_int lda, double anorm, double *rcond) {
  double t[MINI_NMAX * MINI_NMAX], inv[MINI_NMAX * MINI_NMAX];
  int piv[MINI_NMAX];
  if (n > MINI_NMAX) return -3;
  *rcond = 0.0;
  if (n == 0) {
    *rcond = 1.0;
    return 0;
  }
  if (anorm == 0.0) return 0;
  for (int i = 0; i < n; i++) {
    piv[i] = i + 1;
    for (int j = 0; j < n; j++)
      t[j * n + i] = (layout == LAPACK_ROW_MAJOR) ? a[i * lda + j] : a[j * lda + i];
  }
  for (int j = 0; j < n; j++)
    if (t[j * n + j] == 0.0) return 0; /* singular U: rcond = 0 */
  for (int j = 0; j < n; j++) {
    double col[MINI_NMAX];
    for (int i = 0; i < n; i++) col[i] = (i == j) ? 1.0 : 0.0;
    lu_solve_colmajor(n, t, n, piv, col);
    for (int i = 0; i < n; i++) inv[j * n + i] = col[i];
  }
  double ainvnm = 0.0;
  int one = (norm == '1' || norm == 'O' || norm == 'o');
  for (int p = 0; p < n; p++) {
    double s = 0.0;
    for (int q = 0; q < n; q++) s += fabs(one ? inv[p * n + q] : inv[q * n + p]);
    if (s > ainvnm) ainvnm = s;
  }
  if (ainvnm != 0.0) *rcond = (1.0 / ainvnm) / anorm;
  return 0;
}
