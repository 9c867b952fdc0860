/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (never part of the product path).
 *
 * Glue that links against the reference's OWN 2D translation units (compiled
 * from where they lie under /root/reference by oracle/Makefile into
 * oracle/_ref/libnlps2d_ref.so) and exposes them through a flat C interface
 * that tests/ and bench.py's reference arm call with ctypes.
 *
 * What is reference code and what is ours:
 *   - deck parsing, mesh/adjacency construction, particle seeding, LME
 *     initialisation and per-step search, N / dN evaluation, kinematics
 *     helpers, constitutive updates: the reference's compiled functions.
 *   - the explicit NPC-FS step loop in refh_step(): OUR restatement of the
 *     *intended* U_Verlet stage order, because the
 *     reference's U_Verlet is non-functional at this snapshot
 *     (Formulations/Displacements/U-Verlet.c:100,173,224: __mass_NODES
 *     always returns EXIT_FAILURE; :137-142 force assembly commented out).
 *     Every arithmetic operation in it is a call into reference code or a
 *     line-by-line restatement of the cited reference loop; internal forces
 *     use the maintained Kirchhoff form of U-Newmark-beta.c:1257-1374.
 *   - the driver's globals (driver-nl-partsol.c:58-71) and the three print
 *     helpers whose TU needs PETSc (InOutFun/print_ScreenMessage.c).
 */
#include <math.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "nl-partsol.h"
#include "Nodes/Shape-Functions.h"
#include "Nodes/Nodes-Tools.h"
#include "Nodes/LME.h"
#include "Constitutive/Constitutive.h"
#include "Formulations/Courant.h"

/* ---- globals the driver normally defines (driver-nl-partsol.c:58-71) ---- */
char ShapeFunctionGP[MAXC];
char SimulationFile[MAXC];
char Static_conditons[MAXC];
char Formulation[MAXC];
char *TimeIntegrationScheme;
bool Flag_Print_Convergence;
Load gravity_field;
bool Driver_EigenErosion;
bool Driver_EigenSoftening;
bool Petsc_Direct_solver;
bool Petsc_Iterative_solver;
int ResultsTimeStep;

void print_Status(char *Message, int Time) { (void)Message; (void)Time; }
void print_step(int Time, int NumTimeStep, double dt) { (void)Time; (void)NumTimeStep; (void)dt; }
void print_convergence_stats(int Time, int NumTimeStep, int Iter, int MaxIter,
                             double Error0, double Error_total, double Error_relative) {
  (void)Time; (void)NumTimeStep; (void)Iter; (void)MaxIter; (void)Error0;
  (void)Error_total; (void)Error_relative;
}

static Mesh FEM_Mesh;
static Particle MPM_Mesh;
static Time_Int_Params Params;
static int g_ready = 0;

/* nodal work arrays of the last refh_step (full-grid indexing, Nn x 2) */
static double *g_mass = NULL, *g_ddis = NULL, *g_force = NULL, *g_acc = NULL,
              *g_react = NULL;
static double g_dt = 0.0;

int refh_init(const char *deck) {
  int STATUS;
  strcpy(SimulationFile, deck);
  strcpy(Formulation, "-u");
  NumberDOF = NumberDimensions;
  Driver_EigenErosion = false;
  Driver_EigenSoftening = false;
  Flag_Print_Convergence = false;
  Params = Solver_selector__InOutFun__(SimulationFile);
  STATUS = Generate_Gravity_Field__InOutFun__(&gravity_field, SimulationFile, Params);
  if (STATUS == EXIT_FAILURE) return 1;
  FEM_Mesh = GramsBox(SimulationFile, Params);
  STATUS = Generate_One_Phase_Analysis__InOutFun__(&MPM_Mesh, SimulationFile, FEM_Mesh, Params);
  if (STATUS == EXIT_FAILURE) return 2;
  GramsOutputs(SimulationFile);
  /* GramsBox malloc()s ActiveNode and never clears it (Read_GramsBox.c:118), while
   * initialize__LME__ only ever sets entries to true (LME.c:122-141): the reference's
   * first neighbour lists depend on uninitialised heap.  The harness pins the
   * deterministic reading (all false before the first activation pass). */
  memset(FEM_Mesh.ActiveNode, 0, sizeof(bool) * FEM_Mesh.NumNodesMesh);
  initialise_shapefun__MeshTools__(MPM_Mesh, FEM_Mesh);
  size_t nb = (size_t)FEM_Mesh.NumNodesMesh * NumberDimensions * sizeof(double);
  g_mass = (double *)calloc(1, nb);
  g_ddis = (double *)calloc(1, nb);
  g_force = (double *)calloc(1, nb);
  g_acc = (double *)calloc(1, nb);
  g_react = (double *)calloc(1, nb);
  g_ready = 1;
  return 0;
}

void refh_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------ sizes / scalars ------------------------- */
int refh_ndim(void) { return NumberDimensions; }
int refh_num_nodes(void) { return FEM_Mesh.NumNodesMesh; }
int refh_num_elems(void) { return FEM_Mesh.NumElemMesh; }
int refh_num_particles(void) { return MPM_Mesh.NumGP; }
int refh_num_steps(void) { return Params.NumTimeStep; }
double refh_delta_x(void) { return FEM_Mesh.DeltaX; }
double refh_cfl(void) { return Params.CFL; }
double refh_cel(void) { return Params.Cel; }
double refh_dt(void) { return U_DeltaT__SolversLib__(MPM_Mesh, FEM_Mesh.DeltaX, Params); }
double refh_gamma_lme(void) { return gamma_LME; }
double refh_tol_zero_lme(void) { return TOL_zero_LME; }
double refh_tol_wrapper_lme(void) { return TOL_wrapper_LME; }
int refh_max_iter_lme(void) { return max_iter_LME; }
double refh_tol_radial(void) { return TOL_Radial_Returning; }
int refh_maxiter_radial(void) { return Max_Iterations_Radial_Returning; }
double refh_thickness(void) { return Thickness_Plain_Stress; }

/* --------------------------------- mesh --------------------------------- */
void refh_get_coords(double *out) {
  memcpy(out, FEM_Mesh.Coordinates.nV,
         sizeof(double) * FEM_Mesh.NumNodesMesh * NumberDimensions);
}
void refh_get_h_avg(double *out) {
  memcpy(out, FEM_Mesh.h_avg, sizeof(double) * FEM_Mesh.NumNodesMesh);
}
static int chain_len(ChainPtr c) { int n = 0; while (c) { n++; c = c->next; } return n; }
static int chain_copy(ChainPtr c, int *out) { int n = 0; while (c) { out[n++] = c->Idx; c = c->next; } return n; }

/* which: 0 = element connectivity, 1 = NodeNeighbour, 2 = NodalLocality_0 (1 ring),
 * 3 = NodalLocality (2 rings), 4 = particle ListNodes.  Chain (traversal) order. */
static ChainPtr *table_of(int which, int *n) {
  switch (which) {
  case 0: *n = FEM_Mesh.NumElemMesh; return FEM_Mesh.Connectivity;
  case 1: *n = FEM_Mesh.NumNodesMesh; return FEM_Mesh.NodeNeighbour;
  case 2: *n = FEM_Mesh.NumNodesMesh; return FEM_Mesh.NodalLocality_0;
  case 3: *n = FEM_Mesh.NumNodesMesh; return FEM_Mesh.NodalLocality;
  case 4: *n = MPM_Mesh.NumGP; return MPM_Mesh.ListNodes;
  }
  *n = 0;
  return NULL;
}
int refh_table_total(int which) {
  int n, tot = 0;
  ChainPtr *t = table_of(which, &n);
  for (int i = 0; i < n; i++) tot += chain_len(t[i]);
  return tot;
}
void refh_table_csr(int which, int *ptr, int *idx) {
  int n, tot = 0;
  ChainPtr *t = table_of(which, &n);
  ptr[0] = 0;
  for (int i = 0; i < n; i++) {
    tot += chain_copy(t[i], idx + tot);
    ptr[i + 1] = tot;
  }
}
void refh_get_active(unsigned char *out) {
  for (int i = 0; i < FEM_Mesh.NumNodesMesh; i++) out[i] = FEM_Mesh.ActiveNode[i] ? 1 : 0;
}

/* ------------------------ boundary conditions / loads ------------------- */
int refh_num_bounds(void) { return FEM_Mesh.Bounds.NumBounds; }
int refh_bound_num_nodes(int b) { return FEM_Mesh.Bounds.BCC_i[b].NumNodes; }
int refh_bound_dim(int b) { return FEM_Mesh.Bounds.BCC_i[b].Dim; }
void refh_bound_nodes(int b, int *out) {
  memcpy(out, FEM_Mesh.Bounds.BCC_i[b].Nodes, sizeof(int) * FEM_Mesh.Bounds.BCC_i[b].NumNodes);
}
/* dir: Dim x NumTimeStep ints; val: Dim x NumTimeStep doubles (0 where inactive) */
void refh_bound_table(int b, int *dir, double *val) {
  Load L = FEM_Mesh.Bounds.BCC_i[b];
  int N = Params.NumTimeStep;
  for (int k = 0; k < L.Dim; k++)
    for (int s = 0; s < N; s++) {
      dir[k * N + s] = L.Dir[k * N + s];
      val[k * N + s] = (L.Dir[k * N + s] == 1) ? L.Value[k].Fx[s] : 0.0;
    }
}
/* gravity table: Ndim x NumTimeStep.  Semantics of the maintained scheme
 * (U-Newmark-beta.c:1539-1543): b[i] = gravity_field.Value[i].Fx[step] when
 * gravity_field.STATUS.  (The stale __gravity_NODES of U-Verlet.c:257-297 reads a
 * Dir table that Read_Generate_Gravity_Field.c:137-148 only fills at [0],[1].) */
void refh_gravity_table(double *g) {
  int N = Params.NumTimeStep;
  for (int k = 0; k < NumberDimensions; k++)
    for (int s = 0; s < N; s++)
      g[k * N + s] = (gravity_field.STATUS == true) ? gravity_field.Value[k].Fx[s] : 0.0;
}
int refh_num_neumann(void) { return MPM_Mesh.Neumann_Contours.NumBounds; }
/* Neumann sets (Particle.Neumann_Contours, NLPS-Read-u-Neumann-Boundary-Conditions.c:46-210): loaded particles and the
 * Dim x NumTimeStep direction / value tables, as refh_bound_* give them for the Dirichlet sets */
int refh_neumann_num_nodes(int b) { return MPM_Mesh.Neumann_Contours.BCC_i[b].NumNodes; }
int refh_neumann_dim(int b) { return MPM_Mesh.Neumann_Contours.BCC_i[b].Dim; }
void refh_neumann_nodes(int b, int *out) {
  memcpy(out, MPM_Mesh.Neumann_Contours.BCC_i[b].Nodes, sizeof(int) * MPM_Mesh.Neumann_Contours.BCC_i[b].NumNodes);
}
void refh_neumann_table(int b, int *dir, double *val) {
  Load L = MPM_Mesh.Neumann_Contours.BCC_i[b];
  int N = Params.NumTimeStep;
  for (int k = 0; k < L.Dim; k++)
    for (int s = 0; s < N; s++) {
      dir[k * N + s] = L.Dir[k * N + s];
      val[k * N + s] = (L.Dir[k * N + s] == 1) ? L.Value[k].Fx[s] : 0.0;
    }
}

/* -------------------------------- materials ----------------------------- */
int refh_num_materials(void) { return MPM_Mesh.NumberMaterials; }
const char *refh_material_type(int m) { return MPM_Mesh.Mat[m].Type; }
/* out[16]: rho,E,nu,ReferencePressure,kappa_0,Hardening_modulus,Plastic_Strain_0,
 * phi,psi,Exponent_Ortiz,Cohesion,alpha_Borja,a1,a2,a3,J2_degradated */
void refh_material_params(int m, double *out) {
  Material M = MPM_Mesh.Mat[m];
  out[0] = M.rho; out[1] = M.E; out[2] = M.nu; out[3] = M.ReferencePressure;
  out[4] = M.kappa_0; out[5] = M.Hardening_modulus; out[6] = M.Plastic_Strain_0;
  out[7] = M.phi_Frictional; out[8] = M.psi_Frictional; out[9] = M.Exponent_Hardening_Ortiz;
  out[10] = M.Cohesion; out[11] = M.alpha_Hardening_Borja;
  out[12] = M.a_Hardening_Borja[0]; out[13] = M.a_Hardening_Borja[1];
  out[14] = M.a_Hardening_Borja[2]; out[15] = M.J2_degradated;
}
/* out[4]: the Voce hardening parameters of Von-Mises (theta, K_0, K_inf, delta) */
void refh_material_voce(int m, double *out) {
  Material M = MPM_Mesh.Mat[m];
  out[0] = M.theta_Hardening_Voce; out[1] = M.K_0_Hardening_Voce;
  out[2] = M.K_inf_Hardening_Voce; out[3] = M.delta_Hardening_Voce;
}

/* ------------------------------ particle fields ------------------------- */
static double *field_ptr(const char *name, int *cols) {
  const int d = NumberDimensions, T = (NumberDimensions == 2) ? 5 : 9;
  Fields *P = &MPM_Mesh.Phi;
  *cols = 1;
  if (!strcmp(name, "x_GC")) { *cols = d; return P->x_GC.nV; }
  if (!strcmp(name, "dis")) { *cols = d; return P->dis.nV; }
  if (!strcmp(name, "D_dis")) { *cols = d; return P->D_dis.nV; }
  if (!strcmp(name, "vel")) { *cols = d; return P->vel.nV; }
  if (!strcmp(name, "acc")) { *cols = d; return P->acc.nV; }
  if (!strcmp(name, "F_n")) { *cols = T; return P->F_n.nV; }
  if (!strcmp(name, "F_n1")) { *cols = T; return P->F_n1.nV; }
  if (!strcmp(name, "DF")) { *cols = T; return P->DF.nV; }
  if (!strcmp(name, "b_e_n")) { *cols = T; return P->b_e_n.nV; }
  if (!strcmp(name, "b_e_n1")) { *cols = T; return P->b_e_n1.nV; }
  if (!strcmp(name, "Stress")) { *cols = T; return P->Stress.nV; }
  if (!strcmp(name, "C_ep")) { *cols = d * d; return P->C_ep.nV; }
  if (!strcmp(name, "J_n")) return P->J_n.nV;
  if (!strcmp(name, "J_n1")) return P->J_n1.nV;
  if (!strcmp(name, "mass")) return P->mass.nV;
  if (!strcmp(name, "rho")) return P->rho.nV;
  if (!strcmp(name, "Vol_0")) return P->Vol_0.nV;
  if (!strcmp(name, "W")) return P->W;
  if (!strcmp(name, "EPS_n")) return P->EPS_n;
  if (!strcmp(name, "EPS_n1")) return P->EPS_n1;
  if (!strcmp(name, "Kappa_n")) return P->Kappa_n;
  if (!strcmp(name, "Kappa_n1")) return P->Kappa_n1;
  if (!strcmp(name, "lambda")) { *cols = d; return MPM_Mesh.lambda.nV; }
  if (!strcmp(name, "Beta")) { /* LME: one thermalisation parameter; aLME: the d x d metric (Generate-One-Phase-Analysis.c:192-202) */
    if (!strcmp(ShapeFunctionGP, "aLME")) *cols = d * d;
    return MPM_Mesh.Beta.nV;
  }
  if (!strcmp(name, "Cut_off_Ellipsoid")) {
    if (strcmp(ShapeFunctionGP, "aLME")) return NULL;
    *cols = d * d;
    return MPM_Mesh.Cut_off_Ellipsoid.nV;
  }
  if (!strcmp(name, "Back_stress")) { *cols = 3; return P->Back_stress.nV; }
  return NULL;
}
int refh_field_cols(const char *name) { int c; return field_ptr(name, &c) ? c : -1; }
int refh_get_field(const char *name, double *out) {
  int c; double *p = field_ptr(name, &c);
  if (!p) return 1;
  memcpy(out, p, sizeof(double) * (size_t)MPM_Mesh.NumGP * c);
  return 0;
}
int refh_set_field(const char *name, const double *in) {
  int c; double *p = field_ptr(name, &c);
  if (!p) return 1;
  memcpy(p, in, sizeof(double) * (size_t)MPM_Mesh.NumGP * c);
  return 0;
}
void refh_get_ints(const char *name, int *out) {
  int *src = NULL;
  if (!strcmp(name, "I0")) src = MPM_Mesh.I0;
  else if (!strcmp(name, "Element_p")) src = MPM_Mesh.Element_p;
  else if (!strcmp(name, "NumberNodes")) src = MPM_Mesh.NumberNodes;
  else if (!strcmp(name, "MatIdx")) src = MPM_Mesh.MatIdx;
  if (src) memcpy(out, src, sizeof(int) * MPM_Mesh.NumGP);
}

/* ------------------------- reference stage functions -------------------- */
int refh_local_search(void) { return local_search__MeshTools__(MPM_Mesh, FEM_Mesh); }

/* N (n values) and dN (n x d) for particle p, in ListNodes order. returns n */
int refh_shape(int p, double *N, double *dN) {
  int n = MPM_Mesh.NumberNodes[p];
  Element Nodes_p = nodal_set__Particles__(p, MPM_Mesh.ListNodes[p], n);
  Matrix Np = compute_N__MeshTools__(Nodes_p, MPM_Mesh, FEM_Mesh);
  Matrix dNp = compute_dN__MeshTools__(Nodes_p, MPM_Mesh, FEM_Mesh);
  memcpy(N, Np.nV, sizeof(double) * n);
  memcpy(dN, dNp.nV, sizeof(double) * n * NumberDimensions);
  free__MatrixLib__(Np);
  free__MatrixLib__(dNp);
  free(Nodes_p.Connectivity);
  return n;
}

/* Stress_integration__Constitutive__ over all particles (Constitutive.c:18) */
int refh_stress_all(void) {
  int bad = 0;
  for (int p = 0; p < MPM_Mesh.NumGP; p++) {
    Material M = MPM_Mesh.Mat[MPM_Mesh.MatIdx[p]];
    if (Stress_integration__Constitutive__(p, MPM_Mesh, M) == EXIT_FAILURE) bad++;
  }
  return bad;
}

/*
 * One material point driven through the reference's stress update:
 * state arrays are length T (5 in 2D).  Used to freeze golden vectors from the
 * parameter sets of the reference's stand-alone tests
 * (tests/Constitutive/Drucker-Prager-Backward-Euler.c:377-472).
 */
int refh_stress_point(int p, const double *DF, const double *F_n1, double J_n1,
                      const double *b_e_n, double eps_n, double kappa_n,
                      double *stress, double *b_e_n1, double *eps_n1,
                      double *kappa_n1, double *W, double *C_ep) {
  const int T = (NumberDimensions == 2) ? 5 : 9, d = NumberDimensions;
  Fields *P = &MPM_Mesh.Phi;
  memcpy(P->DF.nM[p], DF, sizeof(double) * T);
  memcpy(P->F_n1.nM[p], F_n1, sizeof(double) * T);
  P->J_n1.nV[p] = J_n1;
  memcpy(P->b_e_n.nM[p], b_e_n, sizeof(double) * T);
  P->EPS_n[p] = eps_n;
  P->Kappa_n[p] = kappa_n;
  Material M = MPM_Mesh.Mat[MPM_Mesh.MatIdx[p]];
  int st = Stress_integration__Constitutive__(p, MPM_Mesh, M);
  memcpy(stress, P->Stress.nM[p], sizeof(double) * T);
  memcpy(b_e_n1, P->b_e_n1.nM[p], sizeof(double) * T);
  *eps_n1 = P->EPS_n1[p];
  *kappa_n1 = P->Kappa_n1[p];
  *W = P->W[p];
  memcpy(C_ep, P->C_ep.nM[p], sizeof(double) * d * d);
  return st;
}

/* ------------------- restated NPC-FS step (Appendix C) ------------------ */
static double now_s(void) {
#ifdef _OPENMP
  return omp_get_wtime();
#else
  struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
#endif
}

/* stage timers of the last step: search, p2g_mass_disp, kin_stress, force, g2p */
static double g_t[5];
void refh_stage_times(double *out) { memcpy(out, g_t, sizeof(g_t)); }

int refh_step(int TimeStep) {
  const int d = NumberDimensions, T = (NumberDimensions == 2) ? 5 : 9;
  const int Np = MPM_Mesh.NumGP, Nn = FEM_Mesh.NumNodesMesh;
  const int NumTimeStep = Params.NumTimeStep;
  const double gamma = 0.5; /* U-Verlet.c:76 */
  Fields *P = &MPM_Mesh.Phi;
  int STATUS = EXIT_SUCCESS;
  double t0;

  /* 1. time step (U-Verlet.c:91) */
  const double dt = U_DeltaT__SolversLib__(MPM_Mesh, FEM_Mesh.DeltaX, Params);
  g_dt = dt;
  DeltaTimeStep = dt;

  /* 2. search (U-Verlet.c:93) -- reference code, OpenMP inside */
  t0 = now_s();
  if (local_search__MeshTools__(MPM_Mesh, FEM_Mesh) == EXIT_FAILURE) return EXIT_FAILURE;
  g_t[0] = now_s() - t0;

  /* 3. masks (U-Verlet.c:94-98).  The nodal arrays below use full-grid
   * indexing (A, not Nodes2Mask[A]); the Mask is still built because
   * get_set_field__MeshTools__ / the DOF test need it. */
  Mask ActiveNodes = get_active_nodes__MeshTools__(FEM_Mesh);
  Mask FreeDofs = get_active_dofs__MeshTools__(ActiveNodes, FEM_Mesh, TimeStep, NumTimeStep);
  const int Nact = ActiveNodes.Nactivenodes;
  double *M = (double *)calloc((size_t)Nact * d, sizeof(double));
  double *dU = (double *)calloc((size_t)Nact * d, sizeof(double));
  double *Fo = (double *)calloc((size_t)Nact * d, sizeof(double));
  double *Ac = (double *)calloc((size_t)Nact * d, sizeof(double));

  /* 4-6. lumped mass, predictor, nodal increment of displacement
   * (U-Verlet.c:166-225, 229-253, 301-367).  Scatter is serial in the
   * reference (Verlet) / inside omp critical (Newmark, U-Newmark-beta.c:582). */
  t0 = now_s();
  for (int p = 0; p < Np; p++)
    for (int i = 0; i < d; i++) {
      int idx = p * d + i;
      P->D_dis.nV[idx] = dt * P->vel.nV[idx] + 0.5 * DSQR(dt) * P->acc.nV[idx];
      P->vel.nV[idx] += (1 - gamma) * dt * P->acc.nV[idx];
    }
#pragma omp parallel for schedule(static)
  for (int p = 0; p < Np; p++) {
    Element Nodes_p = nodal_set__Particles__(p, MPM_Mesh.ListNodes[p], MPM_Mesh.NumberNodes[p]);
    Matrix N_p = compute_N__MeshTools__(Nodes_p, MPM_Mesh, FEM_Mesh);
    double m_p = P->mass.nV[p];
    for (int A = 0; A < Nodes_p.NumberNodes; A++) {
      int A_mask = ActiveNodes.Nodes2Mask[Nodes_p.Connectivity[A]];
      double m_A_p = m_p * N_p.nV[A];
#pragma omp critical
      {
        for (int i = 0; i < d; i++) {
          M[A_mask * d + i] += m_A_p;
          dU[A_mask * d + i] += m_p * N_p.nV[A] * P->D_dis.nV[p * d + i];
        }
      }
    }
    free__MatrixLib__(N_p);
    free(Nodes_p.Connectivity);
  }
  for (int A = 0; A < Nact; A++)
    for (int i = 0; i < d; i++) dU[A * d + i] = dU[A * d + i] / M[A * d + i];

  /* 7. Dirichlet (U-Verlet.c:458-526) */
  for (int b = 0; b < FEM_Mesh.Bounds.NumBounds; b++) {
    Load L = FEM_Mesh.Bounds.BCC_i[b];
    for (int j = 0; j < L.NumNodes; j++) {
      int Id_mask = ActiveNodes.Nodes2Mask[L.Nodes[j]];
      if (Id_mask == -1) continue;
      for (int k = 0; k < L.Dim; k++)
        if (L.Dir[k * NumTimeStep + TimeStep] == 1)
          dU[Id_mask * d + k] = L.Value[k].Fx[TimeStep];
    }
  }
  g_t[1] = now_s() - t0;

  /* 8. local state (U-Verlet.c:530-676): kinematics then stress */
  t0 = now_s();
  int neg_jac = -1;
#pragma omp parallel for schedule(static)
  for (int p = 0; p < Np; p++) {
    unsigned n = MPM_Mesh.NumberNodes[p];
    Element Nodes_p = nodal_set__Particles__(p, MPM_Mesh.ListNodes[p], n);
    double *dU_Ap = (double *)calloc(n * d, sizeof(double));
    get_set_field__MeshTools__(dU_Ap, dU, Nodes_p, ActiveNodes);
    Matrix gradient_p = compute_dN__MeshTools__(Nodes_p, MPM_Mesh, FEM_Mesh);
    update_increment_Deformation_Gradient__Particles__(P->DF.nM[p], dU_Ap, gradient_p.nV, n);
    update_Deformation_Gradient_n1__Particles__(P->F_n1.nM[p], P->F_n.nM[p], P->DF.nM[p]);
    P->J_n1.nV[p] = I3__TensorLib__(P->F_n1.nM[p]);
    if (P->J_n1.nV[p] <= 0.0) neg_jac = p;
    double Delta_J_p = I3__TensorLib__(P->DF.nM[p]);
    P->rho.nV[p] = P->rho.nV[p] / Delta_J_p;
    free(dU_Ap);
    free__MatrixLib__(gradient_p);
    free(Nodes_p.Connectivity);
  }
  if (neg_jac >= 0) {
    fprintf(stderr, "refh_step: negative jacobian in particle %i\n", neg_jac);
    STATUS = EXIT_FAILURE;
  }
  /* stress loop is serial in the reference (U-Verlet.c:645; orphaned omp for
   * in U-Newmark-beta.c:1215) */
  for (int p = 0; p < Np && STATUS == EXIT_SUCCESS; p++) {
    Material MatProp_p = MPM_Mesh.Mat[MPM_Mesh.MatIdx[p]];
    if (Stress_integration__Constitutive__(p, MPM_Mesh, MatProp_p) == EXIT_FAILURE) {
      fprintf(stderr, "refh_step: Stress_integration failed for particle %i\n", p);
      STATUS = EXIT_FAILURE;
    }
  }
  g_t[2] = now_s() - t0;

  /* 9. nodal forces: internal, Kirchhoff form (U-Newmark-beta.c:1257-1374,
   * sign as U-Verlet.c:784) + Neumann tractions (U-Verlet.c:805-902) */
  t0 = now_s();
#pragma omp parallel for schedule(static)
  for (int p = 0; p < Np; p++) {
    int STATUS_p;
    unsigned n = MPM_Mesh.NumberNodes[p];
    double V0_p = P->Vol_0.nV[p];
    Element Nodes_p = nodal_set__Particles__(p, MPM_Mesh.ListNodes[p], n);
    Matrix dN_n = compute_dN__MeshTools__(Nodes_p, MPM_Mesh, FEM_Mesh);
    double *dN_n1 = push_forward_dN__MeshTools__(dN_n.nV, P->DF.nM[p], n, &STATUS_p);
    double *tau = P->Stress.nM[p];
    for (unsigned A = 0; A < n; A++) {
      int A_mask = ActiveNodes.Nodes2Mask[Nodes_p.Connectivity[A]];
      double f[3];
      for (int i = 0; i < d; i++) {
        f[i] = 0.0;
        for (int j = 0; j < d; j++) f[i] += tau[i * d + j] * dN_n1[A * d + j];
      }
#pragma omp critical
      {
        for (int i = 0; i < d; i++) Fo[A_mask * d + i] -= f[i] * V0_p;
      }
    }
    free__MatrixLib__(dN_n);
    free(dN_n1);
    free(Nodes_p.Connectivity);
  }
  for (int i = 0; i < MPM_Mesh.Neumann_Contours.NumBounds; i++) {
    Load L = MPM_Mesh.Neumann_Contours.BCC_i[i];
    double Tn[3] = {0.0, 0.0, 0.0};
    for (int j = 0; j < L.NumNodes; j++) {
      int p = L.Nodes[j];
      double A0_p = P->Vol_0.nV[p] / Thickness_Plain_Stress;
      Element Nodes_p = nodal_set__Particles__(p, MPM_Mesh.ListNodes[p], MPM_Mesh.NumberNodes[p]);
      Matrix N_p = compute_N__MeshTools__(Nodes_p, MPM_Mesh, FEM_Mesh);
      for (int k = 0; k < d; k++)
        if (L.Dir[k * NumTimeStep + TimeStep] == 1) Tn[k] = L.Value[k].Fx[TimeStep];
      for (int A = 0; A < Nodes_p.NumberNodes; A++) {
        int A_mask = ActiveNodes.Nodes2Mask[Nodes_p.Connectivity[A]];
        for (int k = 0; k < d; k++) Fo[A_mask * d + k] += N_p.nV[A] * Tn[k] * A0_p;
      }
      free__MatrixLib__(N_p);
      free(Nodes_p.Connectivity);
    }
  }
  g_t[3] = now_s() - t0;

  /* 10. nodal equilibrium + G2P (U-Verlet.c:906-1020; gravity as U-Newmark-beta.c:1539) */
  t0 = now_s();
  double b[3] = {0.0, 0.0, 0.0};
  if (gravity_field.STATUS == true) /* U-Newmark-beta.c:1539-1543 */
    for (int k = 0; k < d; k++) b[k] = gravity_field.Value[k].Fx[TimeStep];
  memset(g_react, 0, sizeof(double) * Nn * d);
  double *Re = (double *)calloc((size_t)Nact * d, sizeof(double));
  for (int A = 0; A < Nact; A++)
    for (int i = 0; i < d; i++) {
      if (FreeDofs.Nodes2Mask[A * NumberDOF + i] != -1) {
        Ac[A * d + i] = b[i] + Fo[A * d + i] / M[A * d + i];
      } else {
        Ac[A * d + i] = 0.0;
        Re[A * d + i] = Fo[A * d + i];
      }
    }
#pragma omp parallel for schedule(static)
  for (int p = 0; p < Np; p++) {
    for (int i = 0; i < d; i++) {
      P->acc.nM[p][i] = 0.0;
      P->D_dis.nM[p][i] = 0.0;
    }
    Element Nodes_p = nodal_set__Particles__(p, MPM_Mesh.ListNodes[p], MPM_Mesh.NumberNodes[p]);
    Matrix N_p = compute_N__MeshTools__(Nodes_p, MPM_Mesh, FEM_Mesh);
    for (int A = 0; A < Nodes_p.NumberNodes; A++) {
      int A_mask = ActiveNodes.Nodes2Mask[Nodes_p.Connectivity[A]];
      for (int i = 0; i < d; i++) {
        P->acc.nM[p][i] += N_p.nV[A] * Ac[A_mask * d + i];
        P->D_dis.nM[p][i] += N_p.nV[A] * dU[A_mask * d + i];
      }
    }
    free__MatrixLib__(N_p);
    free(Nodes_p.Connectivity);
  }

  /* 11. corrector + roll (U-Verlet.c:1024-1084) */
  for (int p = 0; p < Np; p++) {
    P->J_n.nV[p] = P->J_n1.nV[p];
    P->Kappa_n[p] = P->Kappa_n1[p];
    P->EPS_n[p] = P->EPS_n1[p];
    for (int i = 0; i < T; i++) P->b_e_n.nM[p][i] = P->b_e_n1.nM[p][i];
    for (int i = 0; i < d; i++) {
      P->vel.nM[p][i] += gamma * dt * P->acc.nM[p][i];
      P->x_GC.nM[p][i] += P->D_dis.nM[p][i];
      P->dis.nM[p][i] += P->D_dis.nM[p][i];
    }
    for (int i = 0; i < T; i++) P->F_n.nM[p][i] = P->F_n1.nM[p][i];
  }
  g_t[4] = now_s() - t0;

  /* export nodal arrays in full-grid indexing (0 on inactive nodes) */
  memset(g_mass, 0, sizeof(double) * Nn * d);
  memset(g_ddis, 0, sizeof(double) * Nn * d);
  memset(g_force, 0, sizeof(double) * Nn * d);
  memset(g_acc, 0, sizeof(double) * Nn * d);
  for (int A = 0; A < Nn; A++) {
    int Am = ActiveNodes.Nodes2Mask[A];
    if (Am < 0) continue;
    for (int i = 0; i < d; i++) {
      g_mass[A * d + i] = M[Am * d + i];
      g_ddis[A * d + i] = dU[Am * d + i];
      g_force[A * d + i] = Fo[Am * d + i];
      g_acc[A * d + i] = Ac[Am * d + i];
      g_react[A * d + i] = Re[Am * d + i];
    }
  }
  free(M); free(dU); free(Fo); free(Ac); free(Re);
  free(ActiveNodes.Nodes2Mask);
  free(FreeDofs.Nodes2Mask);
  return STATUS;
}

/* which: 0 mass, 1 dU, 2 force, 3 acc, 4 reactions; Nn x d, full-grid indexing */
void refh_get_nodal(int which, double *out) {
  double *src[5] = {g_mass, g_ddis, g_force, g_acc, g_react};
  memcpy(out, src[which], sizeof(double) * FEM_Mesh.NumNodesMesh * NumberDimensions);
}

/* masks as the reference builds them (Nodes-Tools.c:46-156) */
int refh_masks(int step, int *nodes2mask, int *dofs2mask) {
  Mask A = get_active_nodes__MeshTools__(FEM_Mesh);
  Mask D = get_active_dofs__MeshTools__(A, FEM_Mesh, step, Params.NumTimeStep);
  memcpy(nodes2mask, A.Nodes2Mask, sizeof(int) * FEM_Mesh.NumNodesMesh);
  memcpy(dofs2mask, D.Nodes2Mask, sizeof(int) * A.Nactivenodes * NumberDOF);
  int n = A.Nactivenodes;
  free(A.Nodes2Mask);
  free(D.Nodes2Mask);
  return n;
}

/* The reference's own tangent blocks (implicit scheme, K5), callable without a deck: pins the restatement in
 * nlps_oracle.c.  compute_stiffness_elastoplastic__Constitutive__ (Elastoplastic-Tangent-Matrix.c:42-160) and
 * compute_stiffness_density_Neo_Hookean (Neo-Hookean.c:89-141). */
#include "Constitutive/Plasticity/Elastoplastic-Tangent-Matrix.h"
#include "Constitutive/Hyperelastic/Neo-Hookean.h"
int refh_stiffness_ep(double *out, const double *dN_alpha_n1, const double *dN_beta_n1, double *b_e, double *stress,
                      double *C_ep) {
  State_Parameters S;
  memset(&S, 0, sizeof(S));
  S.b_e = b_e;
  S.Stress = stress;
  S.C_ep = C_ep;
  return compute_stiffness_elastoplastic__Constitutive__(out, dN_alpha_n1, dN_beta_n1, S);
}
int refh_stiffness_nh(double *out, const double *dN_alpha_n1, const double *dN_beta_n1, const double *dN_alpha_n,
                      const double *dN_beta_n, double *F_n, double J, double E, double nu) {
  State_Parameters S;
  Material M;
  memset(&S, 0, sizeof(S));
  memset(&M, 0, sizeof(M));
  S.D_phi_n = F_n;
  S.J = J;
  M.E = E;
  M.nu = nu;
  return compute_stiffness_density_Neo_Hookean(out, dN_alpha_n1, dN_beta_n1, dN_alpha_n, dN_beta_n, S, M);
}

/* The particle VTK writers on the harness' current state: the reference's own ASCII writer (binary == 0) or its binary
 * twin of nl-partsol_b200/host/b200_vtk_binary.h (binary == 1; returns 1 when it declines).  Files go to OutputDir. */
#include "b200_vtk_binary.h"
int refh_write_vtk(int step, int results_every, int binary, const char *dir, const char *stem) {
  strcpy(OutputDir, dir);
  strcpy(OutputParticlesFile, stem);
  if (binary) return b200_particle_results_vtk_binary(MPM_Mesh, step, results_every);
  particle_results_vtk__InOutFun__(MPM_Mesh, step, results_every);
  return 0;
}
void refh_set_outputs(int all) {
  Out_global_coordinates = Out_mass = Out_density = Out_nodal_idx = Out_material_idx = Out_velocity = Out_acceleration =
      Out_displacement = Out_stress = Out_volumetric_stress = Out_deformation_gradient = Out_energy = Out_EPS = all != 0;
}
