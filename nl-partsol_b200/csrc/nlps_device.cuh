// nlps_device.cuh -- device-side small algebra and constitutive laws (sm_100a).
//
// Register-resident closed forms replacing the reference's heap Matrix/Tensor
// containers and per-particle LAPACK calls (Matlib/*.c, SURVEY section 2 "adjacent").
// Reference citations are relative to nl-partsol/src.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define NLPS_TOL_NR 10E-6                 // Macros.h:40
#define NLPS_PI 3.14159265358979323846    // Macros.h:42

struct MatParams {
  int type;
  double rho, E, nu, p_ref, kappa_0, H, eps_0, phi, psi, m_exp, cohesion, alpha_borja, a1, a2, a3;
  // constants of the material, computed ONCE on the host with the C library the reference itself uses (mat_hoist):
  // no tan / sin / cos / sqrt per particle per step, and the values are bit-identical to the reference's
  double K, G, lame;                               // bulk, shear and Lame moduli
  double dp_alpha_F, dp_alpha_Q, dp_beta, dp_ads;  // Drucker-Prager cone constants (plane strain or 3D) and sqrt(1 + 3 alpha_Q^2)
  double mn_c;                                     // Matsuoka-Nakai: cohesion / tan(phi)
  double voce_theta, voce_K0, voce_Kinf, voce_delta;  // Von-Mises: theta / K_0 / K_inf / delta _Hardening_Voce (Types.h Material)
};
// Drucker-Prager.c:361-375 (cone constants), Neo-Hookean.c:17-35, Matsuoka-Nakai.c:320-340 (elastic constants)
static inline void mat_hoist(MatParams& m, int ndim) {
  m.K = m.E / (3.0 * (1.0 - 2.0 * m.nu));
  m.G = m.E / (2.0 * (1.0 + m.nu));
  m.lame = m.E * m.nu / ((1.0 + m.nu) * (1.0 - 2.0 * m.nu));
  const double rphi = (NLPS_PI / 180.0) * m.phi, rpsi = (NLPS_PI / 180.0) * m.psi;
  if (ndim == 2) {  // plane-strain cone constants :361-368
    const double tp = tan(rphi), tq = tan(rpsi);
    m.dp_alpha_F = sqrt(2. / 3.) * tp / sqrt(3. + 4. * (tp * tp));
    m.dp_alpha_Q = sqrt(2. / 3.) * tq / sqrt(3. + 4. * (tq * tq));
    m.dp_beta = sqrt(2. / 3.) * 3. / sqrt(3. + 4. * (tp * tp));
  } else {  // :370-375
    m.dp_alpha_F = sqrt(2 / 3.) * 2 * sin(rphi) / (3 - sin(rphi));
    m.dp_alpha_Q = sqrt(2 / 3.) * 2 * sin(rpsi) / (3 - sin(rpsi));
    m.dp_beta = sqrt(2 / 3.) * 6 * cos(rphi) / (3 - sin(rphi));
  }
  m.dp_ads = sqrt(1.0 + 3.0 * m.dp_alpha_Q * m.dp_alpha_Q);
  m.mn_c = rphi > 0.0 ? m.cohesion / tan(rphi) : 0.0;
}

struct ReturnMapParams {
  double tol;
  int max_iter;
  int quirk_rows;  // plastic branches index the eigenvector matrix by ROW (F10-i)
  int want_cep;
};

// ---------------------------------------------------------------------------
// d x d helpers (row-major), D is a compile-time constant
template <int D>
__device__ __forceinline__ double det(const double* A) {
  if (D == 2) return A[0] * A[3] - A[1] * A[2];  // TensorLib.c:154-168
  return A[0] * A[4] * A[8] - A[0] * A[5] * A[7] + A[1] * A[5] * A[6] - A[1] * A[3] * A[8] +
         A[2] * A[3] * A[7] - A[2] * A[4] * A[6];
}

// inverse by cofactors; returns det (caller checks)
template <int D>
__device__ __forceinline__ double inverse(const double* A, double* Ai) {
  double dt = det<D>(A);
  double r = 1.0 / dt;
  if (D == 2) {
    Ai[0] = A[3] * r;
    Ai[1] = -A[1] * r;
    Ai[2] = -A[2] * r;
    Ai[3] = A[0] * r;
  } else {
    Ai[0] = (A[4] * A[8] - A[5] * A[7]) * r;
    Ai[1] = (A[2] * A[7] - A[1] * A[8]) * r;
    Ai[2] = (A[1] * A[5] - A[2] * A[4]) * r;
    Ai[3] = (A[5] * A[6] - A[3] * A[8]) * r;
    Ai[4] = (A[0] * A[8] - A[2] * A[6]) * r;
    Ai[5] = (A[2] * A[3] - A[0] * A[5]) * r;
    Ai[6] = (A[3] * A[7] - A[4] * A[6]) * r;
    Ai[7] = (A[1] * A[6] - A[0] * A[7]) * r;
    Ai[8] = (A[0] * A[4] - A[1] * A[3]) * r;
  }
  return dt;
}

// rcond as the reference computes it: LAPACKE_dgecon on the UNFACTORISED matrix
// (TensorLib.c:966-993): the entries are read as unit-lower L and upper U factors;
// rcond = 1 / (||A||_1 * ||inv(L U)||_1).  Exact norm instead of the Hager estimate.
template <int D>
__device__ inline double rcond_as_reference(const double* A) {
  double anorm = 0.0;
#pragma unroll
  for (int j = 0; j < D; j++) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) s += fabs(A[i * D + j]);
    anorm = fmax(anorm, s);
  }
  if (anorm == 0.0) return 0.0;
#pragma unroll
  for (int j = 0; j < D; j++)
    if (A[j * D + j] == 0.0) return 0.0;
  double ainv = 0.0;
#pragma unroll
  for (int c = 0; c < D; c++) {  // column c of inv(L U)
    double y[D];
#pragma unroll
    for (int i = 0; i < D; i++) {
      double v = (i == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < D; k++)
        if (k < i) v -= A[i * D + k] * y[k];
      y[i] = v;
    }
    double s = 0.0;
#pragma unroll
    for (int i = D - 1; i >= 0; i--) {
      double v = y[i];
#pragma unroll
      for (int k = 0; k < D; k++)
        if (k > i) v -= A[i * D + k] * y[k];
      y[i] = v / A[i * D + i];
      s += fabs(y[i]);
    }
    ainv = fmax(ainv, s);
  }
  return (1.0 / ainv) / anorm;
}

// ---------------------------------------------------------------------------
// DSYEV('V','U') for 2x2 exactly as DSYTRD(=identity)+DSTEQR(->DLAEV2)+sort
// (LAPACK, version unpinned by the reference; call site Drucker-Prager.c:635).
// Signs matter because of F10-i.  w ascending, z row-major, eigenvector j = column j.
__device__ inline void dlaev2_dev(double a, double b, double c, double& rt1, double& rt2, double& cs1,
                                  double& sn1) {
  double sm = a + c, df = a - c, adf = fabs(df), tb = b + b, ab = fabs(tb);
  double acmx, acmn, rt;
  int sgn1, sgn2;
  if (fabs(a) > fabs(c)) { acmx = a; acmn = c; } else { acmx = c; acmn = a; }
  if (adf > ab) { double q = ab / adf; rt = adf * sqrt(1.0 + q * q); }
  else if (adf < ab) { double q = adf / ab; rt = ab * sqrt(1.0 + q * q); }
  else rt = ab * sqrt(2.0);
  if (sm < 0.0) {
    rt1 = 0.5 * (sm - rt); sgn1 = -1;
    rt2 = __dsub_rn(__dmul_rn(__ddiv_rn(acmx, rt1), acmn), __dmul_rn(__ddiv_rn(b, rt1), b));
  } else if (sm > 0.0) {
    rt1 = 0.5 * (sm + rt); sgn1 = 1;
    rt2 = __dsub_rn(__dmul_rn(__ddiv_rn(acmx, rt1), acmn), __dmul_rn(__ddiv_rn(b, rt1), b));
  } else { rt1 = 0.5 * rt; rt2 = -0.5 * rt; sgn1 = 1; }
  double cs;
  if (df >= 0.0) { cs = df + rt; sgn2 = 1; } else { cs = df - rt; sgn2 = -1; }
  if (fabs(cs) > ab) {
    double ct = -tb / cs;
    sn1 = 1.0 / sqrt(1.0 + ct * ct);
    cs1 = ct * sn1;
  } else if (ab == 0.0) { cs1 = 1.0; sn1 = 0.0; }
  else {
    double tn = -cs / tb;
    cs1 = 1.0 / sqrt(1.0 + tn * tn);
    sn1 = tn * cs1;
  }
  if (sgn1 == sgn2) { double tn = cs1; cs1 = -sn1; sn1 = tn; }
}

__device__ inline void dsyev2_dev(double d1, double e, double d2, double* w, double* z) {
  const double eps = 1.1102230246251565e-16;  // 2^-53, dlamch('E')
  const double safmin = 2.2250738585072014e-308;
  double z11 = 1.0, z12 = 0.0, z21 = 0.0, z22 = 1.0;
  double tst = fabs(e);
  bool split = (tst == 0.0) || (tst <= (sqrt(fabs(d1)) * sqrt(fabs(d2))) * eps);
  if (!split) split = (tst * tst <= (eps * eps * fabs(d1)) * fabs(d2) + safmin);
  if (!split) {
    double rt1, rt2, c, s;
    dlaev2_dev(d1, e, d2, rt1, rt2, c, s);
    z11 = c; z12 = -s; z21 = s; z22 = c;
    d1 = rt1; d2 = rt2;
  }
  if (d2 < d1) {
    double t = d1; d1 = d2; d2 = t;
    t = z11; z11 = z12; z12 = t;
    t = z21; z21 = z22; z22 = t;
  }
  w[0] = d1; w[1] = d2;
  z[0] = z11; z[1] = z12; z[2] = z21; z[3] = z22;
}

// cyclic Jacobi for symmetric 3x3 (3D has no compilable reference; convention pinned
// by oracle/mini_lapack.c nlps_jacobi_eig: ascending, largest-|.| component positive)
__device__ inline void jacobi3_dev(const double* a_in, double* w, double* z) {
  double a[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      a[i * 3 + j] = (j >= i) ? a_in[i * 3 + j] : a_in[j * 3 + i];
      z[i * 3 + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
    double diag = a[0] * a[0] + a[4] * a[4] + a[8] * a[8];
    if (off <= 1e-34 * diag || off == 0.0) break;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = i + 1; j < 3; j++) {
        double apq = a[i * 3 + j];
        if (apq == 0.0) continue;
        double theta = (a[j * 3 + j] - a[i * 3 + i]) / (2.0 * apq);
        double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          double akp = a[k * 3 + i], akq = a[k * 3 + j];
          a[k * 3 + i] = c * akp - s * akq;
          a[k * 3 + j] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          double apk = a[i * 3 + k], aqk = a[j * 3 + k];
          a[i * 3 + k] = c * apk - s * aqk;
          a[j * 3 + k] = s * apk + c * aqk;
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          double zkp = z[k * 3 + i], zkq = z[k * 3 + j];
          z[k * 3 + i] = c * zkp - s * zkq;
          z[k * 3 + j] = s * zkp + c * zkq;
        }
      }
  }
  w[0] = a[0]; w[1] = a[4]; w[2] = a[8];
#pragma unroll
  for (int i = 0; i < 2; i++) {
    int k = i;
#pragma unroll
    for (int j = i + 1; j < 3; j++)
      if (w[j] < w[k]) k = j;
    if (k != i) {
      double t = w[i]; w[i] = w[k]; w[k] = t;
#pragma unroll
      for (int j = 0; j < 3; j++) { t = z[j * 3 + i]; z[j * 3 + i] = z[j * 3 + k]; z[j * 3 + k] = t; }
    }
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    int kmax = 0;
    if (fabs(z[3 + j]) > fabs(z[kmax * 3 + j])) kmax = 1;
    if (fabs(z[6 + j]) > fabs(z[kmax * 3 + j])) kmax = 2;
    if (z[kmax * 3 + j] < 0.0) { z[j] = -z[j]; z[3 + j] = -z[3 + j]; z[6 + j] = -z[6 + j]; }
  }
}

// ---------------------------------------------------------------------------
// Constitutive laws.  Tensor storage: 2D T=5 (in-plane 2x2 row-major + slot 4 = 33),
// 3D T=9 (SURVEY Appendix B, U-Analisys.c:22-42).

// Neo-Hookean (Wriggers): tau = lambda/2 (J^2-1) I + G (b - I)   Neo-Hookean.c:17-85
template <int D>
__device__ inline void stress_neo_hookean(const MatParams& m, const double* F, double J, double* tau,
                                          double& W) {
  const double G = m.G, lam = m.lame;
  double c0 = lam * 0.5 * (J * J - 1.0);
  double I1 = 0.0;
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = 0; j < D; j++) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++) s += F[i * D + k] * F[j * D + k];
      double id = (i == j) ? 1.0 : 0.0;
      tau[i * D + j] = c0 * id + G * (s - id);
      if (i == j) I1 += s;
    }
  if (D == 2) tau[4] = c0;
  double lJ = log(J);
  W = 0.25 * lam * (J * J - 1) - 0.5 * lam * lJ - G * lJ + 0.5 * G * (I1 - D);
}

// trial b_e = DF b_e DF^T and its spectral decomposition
// (Drucker-Prager.c:617-661, Matsuoka-Nakai.c:705-746)
template <int D>
__device__ inline void trial_be(const double* be, const double* dphi, double* eval, double* evec) {
  double bt[D * D];
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = 0; j < D; j++) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++)
#pragma unroll
        for (int l = 0; l < D; l++) s += dphi[i * D + k] * be[k * D + l] * dphi[j * D + l];
      bt[i * D + j] = s;
    }
  if (D == 2) {
    dsyev2_dev(bt[0], bt[1], bt[3], eval, evec);
    eval[2] = be[4];
  } else {
    jacobi3_dev(bt, eval, evec);
  }
}

template <int D>
__device__ inline void spectral_sum(const double* v, const double* evec, bool rows, double* out) {
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = 0; j < D; j++) {
      double s = 0.0;
#pragma unroll
      for (int A = 0; A < D; A++) {
        double a = rows ? evec[A * D + i] : evec[A + i * D];
        double b = rows ? evec[A * D + j] : evec[A + j * D];
        s += v[A] * a * b;
      }
      out[i * D + j] = s;
    }
}

template <int D>
__device__ inline void corrector_be(double* be, const double* evec, const double* Eh) {
  double ev[3] = {exp(2 * Eh[0]), exp(2 * Eh[1]), exp(2 * Eh[2])};
  spectral_sum<D>(ev, evec, false, be);
  if (D == 2) be[4] = ev[2];
}

// Drucker-Prager, backward Euler radial return, classical + apex branches
// (Drucker-Prager.c:319-613 and helpers :617-1236).  Returns 0 or an error code.
template <int D>
__device__ inline int stress_drucker_prager(const MatParams& m, const ReturnMapParams& rp, const double* dphi,
                                            double* be, double& eps, double& kappa, double* tau, double& W,
                                            double* cep) {
  double eval[3] = {0, 0, 0}, evec[D * D];
  double Eh[3], Tvol[3], Tdev[3], Tp[3] = {0, 0, 0};
  trial_be<D>(be, dphi, eval, evec);
#pragma unroll
  for (int i = 0; i < 3; i++) Eh[i] = 0.5 * log(eval[i]);
  // cone constants (:361-375) and elastic moduli come from the host (mat_hoist)
  const double K = m.K, G = m.G;
  const double alpha_F = m.dp_alpha_F, alpha_Q = m.dp_alpha_Q, beta = m.dp_beta;
  double n[3] = {0, 0, 0}, dEp[3] = {0, 0, 0};
  double PHI, PHI_0, d_PHI, J2, pressure, dg = 0;
  const double eps_n = eps;
  double eps_k = eps_n, kappa_k = kappa, dkappa = 0.0;
  const double TOL = rp.tol;
  int Iter = 0;
  const double ads = m.dp_ads;
  const double trE = Eh[0] + Eh[1] + Eh[2];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    double Ev = (1.0 / 3.0) * trE;
    Tvol[i] = -m.p_ref - K * Ev;
    Tdev[i] = 2 * G * (Eh[i] - Ev);
  }
  pressure = (Tvol[0] + Tvol[1] + Tvol[2]) / 3.0;
  J2 = sqrt(Tdev[0] * Tdev[0] + Tdev[1] * Tdev[1] + Tdev[2] * Tdev[2]);
#define NLPS_YIELD_CL(dgk, kk) \
  (J2 - 2.0 * G * (dgk)-3.0 * alpha_F * (pressure - 3.0 * K * alpha_Q * (dgk)) - beta * (kk))
  PHI = PHI_0 = NLPS_YIELD_CL(dg, kappa_k);
  bool rows = false;
  if (PHI_0 <= NLPS_TOL_NR) {
#pragma unroll
    for (int i = 0; i < 3; i++) Tp[i] = -Tvol[i] + Tdev[i];
    if (rp.want_cep)
#pragma unroll
      for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = 0; j < D; j++) cep[i * D + j] = (1.0 / 3.0) * K + 2.0 * G * ((i == j ? 1.0 : 0.0) - (1.0 / 3.0));
  } else {
    rows = rp.quirk_rows != 0;
    if (J2 > NLPS_TOL_NR) {
#pragma unroll
      for (int i = 0; i < 3; i++) n[i] = Tdev[i] / J2;
    }
    double base = 1.0 + eps_n / m.eps_0;
    if (base < 0.0) return 6;
    dkappa = (m.kappa_0 / (m.m_exp * m.eps_0)) * pow(base, 1.0 / m.m_exp - 1.0);
    if (alpha_F == 0.0) return 6;
    double plim = 3.0 * alpha_Q * K / (2.0 * G) * J2 +
                  beta / (3.0 * alpha_F) * ((J2 / (2.0 * G)) * dkappa * ads + kappa_k);
    if (-pressure < plim) {
      while (fabs(PHI / PHI_0) >= TOL) {
        Iter++;
        if (Iter == rp.max_iter) break;
        d_PHI = +9.0 * K * alpha_F * alpha_Q - 2.0 * G - beta * dkappa * ads;
        if (fabs(d_PHI) < TOL) return 6;
        dg += -PHI / d_PHI;
        if (dg < 0.0) return 6;
        eps_k = eps_n + dg * ads;
        if (eps_k < 0.0) return 6;
        base = 1.0 + eps_k / m.eps_0;
        if (base < 0.0) return 6;
        kappa_k = m.kappa_0 * pow(base, 1.0 / m.m_exp);
        if (kappa_k < 0.0) return 6;
        dkappa = (m.kappa_0 / (m.m_exp * m.eps_0)) * pow(base, 1.0 / m.m_exp - 1.0);
        PHI = NLPS_YIELD_CL(dg, kappa_k);
      }
#pragma unroll
      for (int i = 0; i < 3; i++) {
        Tp[i] = -Tvol[i] + Tdev[i] + dg * (3 * K * alpha_Q - 2 * G * n[i]);
        dEp[i] = dg * (alpha_Q + n[i]);
      }
      eps = eps_k;
      kappa = kappa_k;
      if (rp.want_cep) {
        double c0 = 9 * alpha_F * alpha_Q * K + 2 * G + beta * dkappa * sqrt(2. / 3. * (1 + 3 * alpha_Q * alpha_Q));
        double c1 = 1.0 - 9.0 * alpha_F * alpha_Q * K / c0, c2 = 0.0;
        if (J2 > NLPS_TOL_NR) c2 = dg / J2;
#pragma unroll
        for (int i = 0; i < D; i++)
#pragma unroll
          for (int j = 0; j < D; j++)
            cep[i * D + j] = c1 * K + 2 * G * ((i == j ? 1.0 : 0.0) - (1. / 3.) * (1.0 - 2.0 * G * c2)) -
                             (6.0 * alpha_Q * K * G / c0) * n[j] - (6.0 * alpha_Q * K * G / c0) * n[i] -
                             4 * G * G * (1.0 / c0 - c2) * n[i] * n[j];
      }
    } else {
      double dg1 = J2 / (2.0 * G), dg2 = 0.0;
      dg = dg1 + dg2;
      while (fabs(PHI / PHI_0) >= TOL) {
        Iter++;
        if (Iter == rp.max_iter) break;
        d_PHI = 3.0 * alpha_Q * K + 3.0 * dkappa * beta * (alpha_Q * alpha_Q) * dg /
                                        (3.0 * alpha_F * sqrt((dg1 * dg1) + 3.0 * (alpha_Q * alpha_Q) * (dg * dg)));
        if (fabs(d_PHI) < TOL) break;
        dg2 += -PHI / d_PHI;
        if (dg2 < 0.0) { dg = 0.0; dg2 = 0.0; break; }
        dg = dg1 + dg2;
        PHI = (beta / (3.0 * alpha_F) *
                   (kappa_k + dkappa * sqrt((dg1 * dg1) + 3.0 * (alpha_Q * alpha_Q) * (dg * dg))) -
               pressure + 3.0 * K * alpha_Q * dg);
      }
      eps_k = eps_n + dg * ads;
      if (eps_k < 0.0) return 6;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        Tp[i] = -Tvol[i] + dg * 3 * K * alpha_Q;
        dEp[i] = dg * alpha_Q + dg1 * n[i];
      }
      eps = eps_k;
      kappa = kappa_k;
      if (rp.want_cep) {
        double c0 = 0.0, c1 = 0.0;
        if (dg > 0.0) {
          c0 = (alpha_Q * beta * sqrt(2. / 3.) * dkappa * dg) /
               (3.0 * alpha_F * K * sqrt(dg1 * dg1 + 3.0 * alpha_Q * alpha_Q * dg * dg) +
                alpha_Q * beta * sqrt(2. / 3.) * dkappa * dg);
          c1 = c0 * K / (2.0 * alpha_Q * G * dg);
        }
#pragma unroll
        for (int i = 0; i < D; i++)
#pragma unroll
          for (int j = 0; j < D; j++) cep[i * D + j] = c0 * K + c1 * n[j];
      }
    }
  }
#undef NLPS_YIELD_CL
  spectral_sum<D>(Tp, evec, rows, tau);
  if (D == 2) tau[4] = Tp[2];
#pragma unroll
  for (int i = 0; i < 3; i++) Eh[i] -= dEp[i];
  W = 0.5 * (Tp[0] * Eh[0] + Tp[1] * Eh[1] + Tp[2] * Eh[2]);
  corrector_be<D>(be, evec, Eh);
  return 0;
}

// 5x5 LU with partial pivoting (first maximal pivot, as IDAMAX) + solve, row-major,
// in place (LAPACKE_dgetrf/dgetrs call at Matsuoka-Nakai.c:1122-1166).  Returns 0 / info.
__device__ inline int lu5_solve(double* A, double* b) {
  for (int j = 0; j < 5; j++) {
    int jp = j;
    double amax = fabs(A[j * 5 + j]);
    for (int i = j + 1; i < 5; i++)
      if (fabs(A[i * 5 + j]) > amax) { amax = fabs(A[i * 5 + j]); jp = i; }
    if (A[jp * 5 + j] == 0.0) return j + 1;
    if (jp != j) {
      for (int k = 0; k < 5; k++) { double t = A[j * 5 + k]; A[j * 5 + k] = A[jp * 5 + k]; A[jp * 5 + k] = t; }
      double t = b[j]; b[j] = b[jp]; b[jp] = t;
    }
    double r = 1.0 / A[j * 5 + j];
    for (int i = j + 1; i < 5; i++) {
      double l = A[i * 5 + j] * r;
      A[i * 5 + j] = l;
      for (int k = j + 1; k < 5; k++) A[i * 5 + k] -= l * A[j * 5 + k];
      b[i] -= l * b[j];
    }
  }
  for (int j = 4; j >= 0; j--) {
    double v = b[j];
    for (int k = j + 1; k < 5; k++) v -= A[j * 5 + k] * b[k];
    b[j] = v / A[j * 5 + j];
  }
  return 0;
}

// Matsuoka-Nakai with Borja hardening: monolithic 5x5 Newton + line search
// (Matsuoka-Nakai.c:300-701 and helpers :705-1290)
struct MNPar { double a0, a1, a2, alpha, c, E, nu, Lame, G; };
__device__ __forceinline__ void mn_Eh(const MNPar& q, const double* T, double* Eh) {
  const double c1 = 1.0 / q.E, c2 = -q.nu / q.E;
  double t0 = T[0] + q.c, t1 = T[1] + q.c, t2 = T[2] + q.c;
  Eh[0] = c1 * t0 + c2 * t1 + c2 * t2;
  Eh[1] = c2 * t0 + c1 * t1 + c2 * t2;
  Eh[2] = c2 * t0 + c2 * t1 + c1 * t2;
}
// LD = Lade-Duncan (Lade-Duncan.c:966-1032): the same monolithic return mapping with the yield surface
// cbrt((27 + kappa) I3) - I1 instead of Matsuoka-Nakai's cbrt((9 + kappa) I3) - cbrt(I1 I2)
template <bool LD>
__device__ __forceinline__ double mn_F(double kphi, double I1, double I2, double I3) {
  return LD ? cbrt((27.0 + kphi) * I3) - I1 : cbrt((9.0 + kphi) * I3) - cbrt(I1 * I2);
}
template <bool LD>
__device__ __forceinline__ void mn_dGdS(double* g, const double* T, double I1, double I2, double I3, double kpsi) {
  const double K2 = (LD ? 27.0 : 9.0) + kpsi, ck = cbrt(K2 * I3);
  if (LD) {
#pragma unroll
    for (int i = 0; i < 3; i++) g[i] = ck / (3.0 * T[i]) - 1.0;
  } else {
    const double cb = cbrt(I1 * I2);
#pragma unroll
    for (int i = 0; i < 3; i++) g[i] = ck / (3.0 * T[i]) - (I1 * (I1 - T[i]) + I2) / (3.0 * (cb * cb));
  }
}
__device__ __forceinline__ double mn_residual(double* R, const double* Etr, const double* Ek, const double* dG,
                                              double kap0, double kaphat0, double Fk, double dl) {
#pragma unroll
  for (int i = 0; i < 3; i++) R[i] = Ek[i] - Etr[i] + dl * dG[i];
  R[3] = kap0 - kaphat0;
  R[4] = Fk;
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 5; i++) s += R[i] * R[i];
  return sqrt(s);
}

template <int D, bool LD = false>
__device__ inline int stress_matsuoka_nakai(const MatParams& m, const ReturnMapParams& rp, const double* dphi,
                                            double* be, double& eps, double& kappa, double* tau, double& W,
                                            double* cep) {
  double eval[3] = {0, 0, 0}, evec[D * D];
  double Etr[3], Ek1[3] = {0, 0, 0}, Ek2[3] = {0, 0, 0};
  trial_be<D>(be, dphi, eval, evec);
#pragma unroll
  for (int i = 0; i < 3; i++) Etr[i] = 0.5 * log(eval[i]);
  MNPar q;
  q.E = m.E; q.nu = m.nu;
  q.Lame = m.lame;
  q.G = m.G;
  q.c = m.mn_c;
  q.alpha = m.alpha_borja; q.a0 = m.a1; q.a1 = m.a2; q.a2 = m.a3;
  const double AAd = q.Lame + 2 * q.G, AAo = q.Lame;
  const double CCd = 1.0 / q.E, CCo = -q.nu / q.E;
  double F_k1, F_k2 = 0, F_0, I1, I2, I3;
  const double Lambda_n = eps;
  double Lambda_k1, Lambda_k2, dl1, dl2;
  double Ttr[3], Tk1[3], Tk2[3];
  const double kap_n0 = kappa;
  double kap1, kap2, kaphat;
  double dG[3] = {0, 0, 0}, ddG[9], R1[5] = {0, 0, 0, 0, 0}, R2[5] = {0, 0, 0, 0, 0}, TM[25];
#pragma unroll
  for (int i = 0; i < 9; i++) ddG[i] = 0.0;
  const double TOL = rp.tol, TOL_apex = 0.1;
  double N0, N1, N2 = 0, delta = 1;
  const int MaxIter_k1 = rp.max_iter, MaxIter_k2 = 10 * rp.max_iter;
  int Iter_k1 = 0, Iter_k2 = 0;
  Ttr[0] = AAd * Etr[0] + AAo * Etr[1] + AAo * Etr[2] - q.c;
  Ttr[1] = AAo * Etr[0] + AAd * Etr[1] + AAo * Etr[2] - q.c;
  Ttr[2] = AAo * Etr[0] + AAo * Etr[1] + AAd * Etr[2] - q.c;
  I1 = Ttr[0] + Ttr[1] + Ttr[2];
  I2 = Ttr[0] * Ttr[1] + Ttr[1] * Ttr[2] + Ttr[0] * Ttr[2];
  I3 = Ttr[0] * Ttr[1] * Ttr[2];
  F_0 = mn_F<LD>(kap_n0, I1, I2, I3);
#pragma unroll
  for (int i = 0; i < 3; i++) Tk1[i] = Ttr[i];
  bool rows = false;
  if (F_0 <= NLPS_TOL_NR) {
    if (rp.want_cep)
#pragma unroll
      for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = 0; j < D; j++) cep[i * D + j] = (i == j) ? AAd : AAo;
  } else {
    rows = rp.quirk_rows != 0;
    mn_Eh(q, Tk1, Ek1);
#pragma unroll
    for (int i = 0; i < 3; i++) {  // Matsuoka-Nakai.c:431-433 overwrites the trial strain, Lade-Duncan.c:430-432 the iterate
      if (LD) Ek1[i] = Etr[i]; else Etr[i] = Ek1[i];
    }
    kaphat = q.a0 * Lambda_n * exp(q.a1 * I1) * exp(-q.a2 * Lambda_n);
    mn_dGdS<LD>(dG, Ttr, I1, I2, I3, q.alpha * kap_n0);
    N0 = mn_residual(R1, Etr, Ek1, dG, kap_n0, kaphat, F_0, 0.0);
    kap1 = kap_n0;
    F_k1 = F_0; dl1 = 0.0; Lambda_k1 = Lambda_n; N1 = N0;
    while ((fabs(N1 / N0) >= TOL) && (fabs(F_k1 / F_0) >= TOL)) {
      delta = 1.0;
      double ek = q.a0 * exp(q.a1 * I1) * exp(-q.a2 * Lambda_k1);
      double dkds = q.a1 * Lambda_k1 * ek;
      double dkdl = (1 - q.a2 * Lambda_k1) * ek;
      const double KOFF = LD ? 27.0 : 9.0;
      double K1 = KOFF + kap1, K2 = KOFF + q.alpha * kap1;
      double cb = cbrt(I1 * I2), ck1 = cbrt(K1 * I3), ck2 = cbrt(K2 * I3), cI3 = cbrt(I3);
      double dFds[3], dg[3], ddGk[3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        dg[i] = LD ? 1.0 : (I1 * (I1 - Tk1[i]) + I2) / (3.0 * (cb * cb));
        dFds[i] = ck1 / (3.0 * Tk1[i]) - dg[i];
        double c2 = cbrt(K2);
        ddGk[i] = LD ? cI3 / (3.0 * Tk1[i]) : (cI3 / (3.0 * Tk1[i])) / (3.0 * (c2 * c2));  // (Lade-Duncan.c:1030-1032 as compiled)
      }
      double c1k = cbrt(K1);
      double dFdk = (1.0 / 3.0) * (1.0 / (c1k * c1k)) * cI3;
#pragma unroll
      for (int A = 0; A < 3; A++)
#pragma unroll
        for (int B = 0; B < 3; B++) {
          double ddg = LD ? 0.0 : (1.0 / (cb * cb)) / 3.0 * (3.0 * I1 - Tk1[A] - Tk1[B] - I1 * (A == B)) -
                                      (2.0 / cb) * dg[A] * dg[B];
          ddG[A * 3 + B] = (1.0 / 3.0) * ck2 * (1.0 / (3.0 * Tk1[A] * Tk1[B]) - 1.0 * (A == B) / (Tk1[A] * Tk1[A])) - ddg;
        }
#pragma unroll
      for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) TM[i * 5 + j] = ((i == j) ? CCd : CCo) + dl1 * ddG[i * 3 + j];
        TM[i * 5 + 3] = q.alpha * dl1 * ddGk[i];
        TM[i * 5 + 4] = dG[i];
        TM[15 + i] = -dkds;
        TM[20 + i] = dFds[i];
      }
      TM[18] = 1.0; TM[19] = -dkdl; TM[23] = dFdk; TM[24] = 0.0;
#pragma unroll
      for (int i = 0; i < 5; i++) TM[i * 5 + i] += R1[i];  // "preconditioner" :511-516
      if (lu5_solve(TM, R1) != 0) return 7;
      dl2 = dl1 - delta * R1[4];
      if (Lambda_n + dl2 < 0.0) break;
      Lambda_k2 = Lambda_n + dl2;
#pragma unroll
      for (int i = 0; i < 3; i++) Tk2[i] = Tk1[i] - delta * R1[i];
      kap2 = kap1 - delta * R1[3];
      Iter_k2 = 0;
      if (fabs((Tk2[0] + Tk2[1] + Tk2[2]) / 3.0) < TOL_apex) {
        Lambda_k2 = Lambda_n; kap2 = kap_n0;
        Tk2[0] = Tk2[1] = Tk2[2] = 0.0;
        // reference `break`s here WITHOUT copying k2 -> k1 (:545-553)
        break;
      }
#define NLPS_MN_EVAL2()                                                       \
  I1 = Tk2[0] + Tk2[1] + Tk2[2];                                              \
  I2 = Tk2[0] * Tk2[1] + Tk2[1] * Tk2[2] + Tk2[0] * Tk2[2];                   \
  I3 = Tk2[0] * Tk2[1] * Tk2[2];                                              \
  mn_Eh(q, Tk2, Ek2);                                                         \
  kaphat = q.a0 * Lambda_k2 * exp(q.a1 * I1) * exp(-q.a2 * Lambda_k2);        \
  mn_dGdS<LD>(dG, Tk2, I1, I2, I3, q.alpha * kap2);                           \
  F_k2 = mn_F<LD>(kap2, I1, I2, I3);                                          \
  N2 = mn_residual(R2, Etr, Ek2, dG, kap2, kaphat, F_k2, dl2);
      NLPS_MN_EVAL2();
      while ((fabs(N2 - N1) > TOL) && (fabs(F_k2 / F_0) >= TOL)) {
        delta = (delta * delta) * 0.5 * N1 / (N2 - delta * N1 + N1);
        if ((delta > 1.0) || (delta < 0.0)) break;
        dl2 = dl1 - delta * R2[4];
        if (Lambda_n + dl2 < 0.0) break;
        Lambda_k2 = Lambda_n + dl2;
#pragma unroll
        for (int i = 0; i < 3; i++) Tk2[i] = Tk1[i] - delta * R2[i];
        kap2 = kap1 - delta * R2[3];
        if (fabs((Tk2[0] + Tk2[1] + Tk2[2]) / 3.0) < TOL_apex) {
          Lambda_k2 = Lambda_n; kap2 = kap_n0;
          Tk2[0] = Tk2[1] = Tk2[2] = 0.0;
          break;
        }
        NLPS_MN_EVAL2();
        Iter_k2++;
        if (Iter_k2 == MaxIter_k2) break;
      }
#undef NLPS_MN_EVAL2
#pragma unroll
      for (int i = 0; i < 3; i++) { Tk1[i] = Tk2[i]; Ek1[i] = Ek2[i]; }
      kap1 = kap2;
      Lambda_k1 = Lambda_k2; F_k1 = F_k2; dl1 = dl2;
#pragma unroll
      for (int i = 0; i < 5; i++) R1[i] = R2[i];
      N1 = N2;
      Iter_k1++;
      if (fabs((Tk1[0] + Tk1[1] + Tk1[2]) / 3.0) < TOL_apex) {
        Lambda_k1 = Lambda_n; kap1 = kap_n0;
        Tk1[0] = Tk1[1] = Tk1[2] = 0.0;
        break;
      }
      if (Iter_k1 == MaxIter_k1) break;
    }
    eps = Lambda_k1;
    kappa = kap1;
    if (rp.want_cep) {  // :1243-1290 ; 3D: the reference never writes C_ep (F10-ii), we do
      double Ca[9], Ci[9];
#pragma unroll
      for (int i = 0; i < 9; i++) Ca[i] = (((i % 4) == 0) ? CCd : CCo) + dl1 * ddG[i];
      double dt = inverse<3>(Ca, Ci);
      if (dt == 0.0) return 7;
      if (D == 2) { cep[0] = Ci[0]; cep[1] = Ci[1]; cep[2] = Ci[3]; cep[3] = Ci[4]; }
      else {
#pragma unroll
        for (int i = 0; i < D * D; i++) cep[i] = Ci[i];
      }
    }
  }
  {
    double v[3] = {Tk1[0] + q.c, Tk1[1] + q.c, Tk1[2] + q.c};
    spectral_sum<D>(v, evec, rows, tau);
    if (D == 2) tau[4] = Tk1[2] + q.c;
    W = 0.5 * (v[0] * Etr[0] + v[1] * Etr[1] + v[2] * Etr[2]);
  }
  corrector_be<D>(be, evec, Ek1);  // Ek1 == 0 in the elastic branch (:305,:699): b_e := I
  return 0;
}

// ---------------------------------------------------------------------------
// Von-Mises (J2) with linear + Voce isotropic and linear kinematic hardening, radial return in principal Hencky strains
// (Constitutive/Plasticity/Von-Mises.c:228-391 and helpers :395-757).  Reproduced as compiled: the volumetric part is
// K tr(E)/3 (:553-559); the elastic branch rotates with eigenvectors in columns (:616-617), the plastic branch with rows
// (:717-718, SURVEY F10-i, rp.quirk_rows); `back` = Phi.Back_stress, PRINCIPAL components, updated in place.
// rp.want_cep: the tangent moduli of __tangent_moduli (:207-226, 730-757) in principal space, as compiled -- "K_iso_k",
// "K_kin_k" are the hardening VALUES kappa_k (not their derivatives), theta = 0 while J2 <= TOL_NR (an unstressed point has
// no shear stiffness); in the elastic branch the reference hands an uninitialised kappa_k that only multiplies n (x) n = 0.
template <int D>
__device__ inline int stress_von_mises(const MatParams& m, const ReturnMapParams& rp, const double* dphi, double* be,
                                       double& eps, double* back, double* tau, double& W, double* cep) {
  double eval[3] = {0, 0, 0}, evec[D * D];
  double Eh[3], Tvol[3], Tdev[3], Tp[3];
  trial_be<D>(be, dphi, eval, evec);
#pragma unroll
  for (int i = 0; i < 3; i++) Eh[i] = 0.5 * log(eval[i]);
  const double K = m.K, G = m.G, sigma_y = m.kappa_0, H = m.H, theta = m.voce_theta, dK = m.voce_Kinf - m.voce_K0,
               delta = m.voce_delta;
  const double Tb[3] = {back[0], back[1], back[2]};
  const double eps_n = eps;
  if (eps_n < 0.0) return NLPS_ERR_RETURN_MAP_VM;  // __kappa :646-647
  const double trE = Eh[0] + Eh[1] + Eh[2];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double Ev = (1.0 / 3.0) * trE;
    Tvol[i] = K * Ev;
    Tdev[i] = 2 * G * (Eh[i] - Ev) - Tb[i];
  }
  const double J2 = sqrt(Tdev[0] * Tdev[0] + Tdev[1] * Tdev[1] + Tdev[2] * Tdev[2]);
  const double s23 = sqrt(2. / 3.);
  const double kin_n = (1 - theta) * H * eps_n;
  double iso_k = sigma_y + theta * H * eps_n + dK * (1 - exp(-delta * eps_n)), kin_k = kin_n;
  const double PHI_0 = J2 - s23 * (iso_k + kin_k - kin_n) - 2.0 * G * 0.0;
  double dEp[3] = {0, 0, 0}, n[3] = {0, 0, 0}, dg = 0.0;
  if (PHI_0 <= 0.0) {
#pragma unroll
    for (int i = 0; i < 3; i++) Tp[i] = Tvol[i] + Tdev[i];  // :578-586 (the back stress is not added back)
    spectral_sum<D>(Tp, evec, false, tau);
  } else {
#pragma unroll
    for (int i = 0; i < 3; i++) n[i] = Tdev[i] / J2;
    double PHI = PHI_0, eps_k = eps_n;
    int Iter = 0;
    while (fabs(PHI / PHI_0) >= rp.tol) {
      Iter++;
      if (Iter == rp.max_iter) break;
      if (eps_k < 0.0) return NLPS_ERR_RETURN_MAP_VM;
      const double d_iso = theta * H + delta * dK * exp(-delta * eps_k), d_kin = (1 - theta) * H;
      const double d_PHI = -2.0 * G * (1.0 + (d_iso + d_kin) / (3 * G));
      dg += -PHI / d_PHI;
      eps_k = eps_n + s23 * dg;
      if (eps_k < 0.0) return NLPS_ERR_RETURN_MAP_VM;
      iso_k = sigma_y + theta * H * eps_k + dK * (1 - exp(-delta * eps_k));
      kin_k = (1 - theta) * H * eps_k;
      PHI = J2 - s23 * (iso_k + kin_k - kin_n) - 2.0 * G * dg;
    }
    const double dKkin = kin_k - kin_n;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      Tp[i] = Tvol[i] + Tdev[i] + Tb[i] - dg * 2 * G * n[i];
      dEp[i] = dg * n[i];
      back[i] = Tb[i] + s23 * dKkin * n[i];
    }
    eps = eps_k;
    spectral_sum<D>(Tp, evec, rp.quirk_rows != 0, tau);
  }
  if (D == 2) tau[4] = Tp[2];
#pragma unroll
  for (int i = 0; i < 3; i++) Eh[i] -= dEp[i];
  corrector_be<D>(be, evec, Eh);
  if (rp.want_cep) {
    const double th = J2 > 10E-6 /* TOL_NR, Macros.h:40 */ ? 1.0 - 2.0 * G * dg / J2 : 0.0;
    const double thb = 1.0 / (1.0 + (iso_k + kin_k) / (3.0 * G)) - (1.0 - th);
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
      for (int j = 0; j < D; j++)
        cep[i * D + j] = K + 2.0 * G * th * ((i == j ? 1.0 : 0.0) - (1.0 / 3.0)) - 2.0 * G * thb * n[i] * n[j];
  }
  W = 0.5 * (Tp[0] * Eh[0] + Tp[1] * Eh[1] + Tp[2] * Eh[2]);
  return 0;
}

// Hencky hyperelasticity (Constitutive/Hyperelastic/Hencky.c:30-93): b = F F^T (compute-Strains.c:365-384), principal
// logarithmic strains (0 out of plane in 2D, :41), T = AA E, eigenvectors in columns (:242-284).
template <int D>
__device__ inline void stress_hencky(const MatParams& m, const double* F, double* tau, double& W) {
  double b[D * D], eval[3] = {0, 0, 1.0}, evec[D * D], Eh[3], Tp[3];
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = i; j < D; j++) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++) s += F[i * D + k] * F[j * D + k];
      b[i * D + j] = s;
      b[j * D + i] = s;
    }
  if (D == 2) dsyev2_dev(b[0], b[1], b[3], eval, evec);
  else jacobi3_dev(b, eval, evec);
#pragma unroll
  for (int i = 0; i < 3; i++) Eh[i] = 0.5 * log(eval[i]);
  const double L = m.lame, L2G = m.lame + 2 * m.G;
  Tp[0] = L2G * Eh[0] + L * Eh[1] + L * Eh[2];
  Tp[1] = L * Eh[0] + L2G * Eh[1] + L * Eh[2];
  Tp[2] = L * Eh[0] + L * Eh[1] + L2G * Eh[2];
  spectral_sum<D>(Tp, evec, false, tau);
  if (D == 2) tau[4] = Tp[2];
  W = 0.5 * (Tp[0] * Eh[0] + Tp[1] * Eh[1] + Tp[2] * Eh[2]);
}

// Stress_integration__Constitutive__ (Constitutive.c:18-258) for every law except Neo-Hookean: which fields of the
// history a law reads and writes.  `back` may be nullptr when the cloud holds no Von-Mises particle.
__host__ __device__ __forceinline__ bool mat_has_history(int mtype) {
  return mtype == NLPS_MAT_DRUCKER_PRAGER || mtype == NLPS_MAT_MATSUOKA_NAKAI || mtype == NLPS_MAT_VON_MISES ||
         mtype == NLPS_MAT_LADE_DUNCAN;
}
template <int D>
__device__ inline int stress_with_history(int mtype, const MatParams& m, const ReturnMapParams& rp, const double* DF,
                                          const double* Fn1, double* be, double& eps, double& kap, double* back,
                                          double* tau, double& W, double* cep) {
  if (mtype == NLPS_MAT_DRUCKER_PRAGER) return stress_drucker_prager<D>(m, rp, DF, be, eps, kap, tau, W, cep);
  if (mtype == NLPS_MAT_MATSUOKA_NAKAI) return stress_matsuoka_nakai<D>(m, rp, DF, be, eps, kap, tau, W, cep);
  if (mtype == NLPS_MAT_VON_MISES) return stress_von_mises<D>(m, rp, DF, be, eps, back, tau, W, cep);
  if (mtype == NLPS_MAT_LADE_DUNCAN) return stress_matsuoka_nakai<D, true>(m, rp, DF, be, eps, kap, tau, W, cep);
  stress_hencky<D>(m, Fn1, tau, W);
  return 0;
}
