/*
 * ref_harness3d.c -- TEST INFRASTRUCTURE.  The reference's OWN 3D constitutive code, callable point-wise.
 *
 * The reference's 3D build does not compile as a whole (Matlib/TensorLib.c has errors in every 3D branch), but
 * the translation units of the elastoplastic laws do, and they depend on nothing but LAPACKE:
 *     Constitutive/Plasticity/Drucker-Prager.c, Matsuoka-Nakai.c, Elastoplastic-Tangent-Matrix.c
 * compiled from where they lie WITHOUT -DUSE_PLAINSTRAIN (NumberDimensions == 3, Macros.h:34-36) by `make ref3d`.  This
 * file gives them the two driver globals they read and flat entry points, so that the 3D branches of the oracle port
 * (oracle/nlps_oracle.c) are pinned to compiled reference code for the laws of BASELINE configs[2]-[4] (tests/golden/
 * {dp,mn,nh}_points3d.npz, tests/test_oracle_3d_laws.py).  Neo-Hookean.c and Particles/compute-Strains.c join them further
 * down; the 3D LME has its own harness (ref_harness3d_lme.c).
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
#include "Constitutive/Hyperelastic/Neo-Hookean.h"
#include "Particles.h" /* prototypes of Particles/compute-Strains.c */

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

/* ---- Neo-Hookean in 3D (BASELINE configs[2] and [4]): Constitutive/Hyperelastic/Neo-Hookean.c and Particles/compute-Strains.c
 * compile in 3D; the Kirchhoff stress, its energy and the tangent block only reach left_Cauchy_Green__Particles__ (plain
 * arithmetic).  The TensorLib symbols other functions of those two TUs reference are unreachable from here and abort. */
static void unreachable3d(const char *what) { fprintf(stderr, "ref_harness3d: %s reached\n", what); abort(); }
#define STUB_TENSOR(name, args) Tensor name args { unreachable3d(#name); Tensor t_; memset(&t_, 0, sizeof(t_)); return t_; }
STUB_TENSOR(alloc__TensorLib__, (int o))
STUB_TENSOR(Identity__TensorLib__, (void))
STUB_TENSOR(Inverse__TensorLib__, (Tensor a))
STUB_TENSOR(dyadic_Product__TensorLib__, (Tensor a, Tensor b))
STUB_TENSOR(vector_linear_mapping__TensorLib__, (Tensor a, Tensor b))
STUB_TENSOR(matrix_product_old__TensorLib__, (Tensor a, Tensor b))
void free__TensorLib__(Tensor a) { (void)a; unreachable3d("free__TensorLib__"); }
/* the strain energy of Neo-Hookean.c calls I1__TensorLib__, whose 3D branch is the line that stops TensorLib.c from compiling
 * (TensorLib.c:120 assigns the trace to an undeclared `I3`): restated as the trace it evidently means */
double I1__TensorLib__(const double *a) { return a[0] + a[4] + a[8]; }
double I3__TensorLib__(const double *a) { (void)a; unreachable3d("I3__TensorLib__"); return 0.0; }
double inner_product__TensorLib__(Tensor a, Tensor b) { (void)a; (void)b; unreachable3d("inner_product__TensorLib__"); return 0.0; }
int compute_inverse__TensorLib__(double *o, const double *a) { (void)o; (void)a; unreachable3d("compute_inverse__TensorLib__"); return 1; }
int compute_adjunt__TensorLib__(double *o, const double *a) { (void)o; (void)a; unreachable3d("compute_adjunt__TensorLib__"); return 1; }
void matrix_product__TensorLib__(double *o, const double *a, const double *b) { (void)o; (void)a; (void)b; unreachable3d("matrix_product__TensorLib__"); }

int refh3_stress_nh(double E, double nu, const double *F_n1, double J, double *stress, double *W) {
  Material M;
  State_Parameters S;
  memset(&M, 0, sizeof(M));
  memset(&S, 0, sizeof(S));
  double Dphi[9];
  memcpy(Dphi, F_n1, sizeof(Dphi));
  M.E = E; M.nu = nu;
  S.Stress = stress; S.D_phi_n1 = Dphi; S.J = J; S.W = W;
  return compute_Kirchhoff_Stress_Neo_Hookean__Constitutive__(S, M);
}
int refh3_stiffness_nh(double *out, const double *dN_alpha_n1, const double *dN_beta_n1, const double *dN_alpha_n,
                       const double *dN_beta_n, double *F_n, double J, double E, double nu) {
  State_Parameters S;
  Material M;
  memset(&S, 0, sizeof(S));
  memset(&M, 0, sizeof(M));
  S.D_phi_n = F_n; S.J = J; M.E = E; M.nu = nu;
  return compute_stiffness_density_Neo_Hookean(out, dN_alpha_n1, dN_beta_n1, dN_alpha_n, dN_beta_n, S, M);
}

/* kinematics of one particle (Particles/compute-Strains.c:20-44, 76-105, compiled in 3D): DF = I + sum_A dU_A (x) grad N_A,
 * F_n1 = DF F_n */
void refh3_kinematics(int n, const double *dU, const double *grad, const double *F_n, double *DF, double *F_n1) {
  update_increment_Deformation_Gradient__Particles__(DF, dU, grad, (unsigned)n);
  update_Deformation_Gradient_n1__Particles__(F_n1, F_n, DF);
}
