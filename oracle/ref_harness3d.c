/*
 * ref_harness3d.c -- TEST INFRASTRUCTURE.  The reference's OWN 3D constitutive code, callable point-wise.
 *
 * The reference's 3D build does not compile as a whole (Matlib/TensorLib.c has errors in every 3D branch), but
 * the translation units of the elastoplastic laws do, and they depend on nothing but LAPACKE:
 *     Constitutive/Plasticity/Drucker-Prager.c, Matsuoka-Nakai.c, Elastoplastic-Tangent-Matrix.c
 * compiled from where they lie WITHOUT -DUSE_PLAINSTRAIN (NumberDimensions == 3, Macros.h:34-36) by `make ref3d`.  This
 * file gives them the two driver globals they read and flat entry points, so that the 3D branches of the oracle port
 * (oracle/nlps_oracle.c) are pinned to compiled reference code for the laws of BASELINE configs[3] (tests/golden/
 * points3d_*.npz, tests/test_oracle_3d_laws.py).  LME, kinematics and Neo-Hookean in 3D stay restated (they need TensorLib).
 */
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Macros.h"
#include "Types.h"
#include "Constitutive/Plasticity/Drucker-Prager.h"
#include "Constitutive/Plasticity/Matsuoka-Nakai.h"
#include "Constitutive/Plasticity/Elastoplastic-Tangent-Matrix.h"

#if NumberDimensions != 3
#error "ref_harness3d.c is the 3D build: compile without -DUSE_PLAINSTRAIN"
#endif

double TOL_Radial_Returning;        /* driver globals (Globals.h) read by the return-mapping loops */
int Max_Iterations_Radial_Returning;

int refh3_ndim(void) { return NumberDimensions; }

static Material fill_material(const char *type, const double *p) {
  Material M;
  memset(&M, 0, sizeof(M));
  strcpy(M.Type, type);
  /* parameter order of oracle/ref_harness.c::refh_material_params */
  M.rho = p[0]; M.E = p[1]; M.nu = p[2]; M.ReferencePressure = p[3];
  M.kappa_0 = p[4]; M.Hardening_modulus = p[5]; M.Plastic_Strain_0 = p[6];
  M.phi_Frictional = p[7]; M.psi_Frictional = p[8]; M.Exponent_Hardening_Ortiz = p[9];
  M.Cohesion = p[10]; M.alpha_Hardening_Borja = p[11];
  M.a_Hardening_Borja[0] = p[12]; M.a_Hardening_Borja[1] = p[13]; M.a_Hardening_Borja[2] = p[14];
  M.J2_degradated = p[15];
  return M;
}

/* one material point through the law exactly as Stress_integration__Constitutive__ prepares it (Constitutive.c:144-214):
 * b_e, EPS, Kappa enter as the step-n values and leave as step n+1 */
int refh3_stress_point(const char *type, const double *mat, double tol_radial, int maxiter_radial, const double *DF,
                       const double *F_n1, const double *b_e_n, double eps_n, double kappa_n, double *stress,
                       double *b_e_n1, double *eps_n1, double *kappa_n1, double *W, double *C_ep) {
  TOL_Radial_Returning = tol_radial;
  Max_Iterations_Radial_Returning = maxiter_radial;
  Material M = fill_material(type, mat);
  State_Parameters S;
  memset(&S, 0, sizeof(S));
  double dphi[9], Dphi[9];
  bool failure = false;
  memcpy(dphi, DF, sizeof(dphi));
  memcpy(Dphi, F_n1, sizeof(Dphi));
  memcpy(b_e_n1, b_e_n, sizeof(double) * 9);
  *eps_n1 = eps_n;
  *kappa_n1 = kappa_n;
  S.Particle_Idx = 0;
  S.Stress = stress;
  S.W = W;
  S.b_e = b_e_n1;
  S.EPS = eps_n1;
  S.Kappa = kappa_n1;
  S.d_phi = dphi;
  S.D_phi_n1 = Dphi;
  S.Failure = &failure;
  S.compute_C_ep = true;
  S.C_ep = C_ep;
  if (!strcmp(type, "Drucker-Prager")) return compute_Kirchhoff_Stress_Drucker_Prager__Constitutive__(S, M);
  if (!strcmp(type, "Matsuoka-Nakai")) return compute_Kirchhoff_Stress_Matsuoka_Nakai__Constitutive__(S, M);
  return -1;
}

/* compute_stiffness_elastoplastic__Constitutive__ (Elastoplastic-Tangent-Matrix.c:42-160), 3D */
int refh3_stiffness_ep(double *out, const double *dN_alpha_n1, const double *dN_beta_n1, double *b_e, double *stress,
                       double *C_ep) {
  State_Parameters S;
  memset(&S, 0, sizeof(S));
  S.b_e = b_e;
  S.Stress = stress;
  S.C_ep = C_ep;
  return compute_stiffness_elastoplastic__Constitutive__(out, dN_alpha_n1, dN_beta_n1, S);
}
