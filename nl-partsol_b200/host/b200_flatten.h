/*
 * b200_flatten.h -- host glue shared by the scheme shims (U-Verlet-b200.c, U-Newmark-beta-b200.c): the reference's
 * `Mesh` / `Particle` / `Time_Int_Params` structs and process globals -> the PODs of include/nlps_b200.h.
 * Compiled against the reference's own headers.  Nothing here is on the stepped path.
 */
#ifndef B200_FLATTEN_H
#define B200_FLATTEN_H
#include <math.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Macros.h"
#include "Types.h"
#include "Globals.h"
#include "Matlib.h"
#include "Particles.h"
#include "InOutFun.h"

#include "nlps_b200.h"

static int chain_len(ChainPtr c) {
  int n = 0;
  while (c) { n++; c = c->next; }
  return n;
}

/* table of chains -> CSR, chain (traversal) order */
static void chains_to_csr(ChainPtr *table, int n, int **ptr, int **idx) {
  int tot = 0;
  *ptr = (int *)malloc(sizeof(int) * (n + 1));
  (*ptr)[0] = 0;
  for (int i = 0; i < n; i++) {
    tot += chain_len(table[i]);
    (*ptr)[i + 1] = tot;
  }
  *idx = (int *)malloc(sizeof(int) * (tot > 0 ? tot : 1));
  for (int i = 0, o = 0; i < n; i++)
    for (ChainPtr c = table[i]; c; c = c->next) (*idx)[o++] = c->Idx;
}

static nlps_load *loads_to_pod(const Boundaries *B, int nsteps) {
  nlps_load *L = (nlps_load *)calloc(B->NumBounds > 0 ? B->NumBounds : 1, sizeof(nlps_load));
  for (int b = 0; b < B->NumBounds; b++) {
    const Load *src = &B->BCC_i[b];
    int *dir = (int *)calloc((size_t)src->Dim * nsteps, sizeof(int));
    double *val = (double *)calloc((size_t)src->Dim * nsteps, sizeof(double));
    for (int k = 0; k < src->Dim; k++)
      for (int s = 0; s < nsteps; s++) {
        dir[k * nsteps + s] = src->Dir[k * nsteps + s];
        if (src->Dir[k * nsteps + s] == 1) val[k * nsteps + s] = src->Value[k].Fx[s];
      }
    L[b].n_ids = src->NumNodes;
    L[b].dim = src->Dim;
    L[b].ids = src->Nodes;
    L[b].dir = dir;
    L[b].val = val;
  }
  return L;
}

static void free_loads(nlps_load *L, int n) {
  for (int b = 0; b < n; b++) {
    free((void *)L[b].dir);
    free((void *)L[b].val);
  }
  free(L);
}

static int material_to_pod(const Material *M, nlps_material *out) {
  memset(out, 0, sizeof(*out));
  if (strcmp(M->Type, "Neo-Hookean-Wriggers") == 0) out->type = NLPS_MAT_NEO_HOOKEAN_WRIGGERS;
  else if (strcmp(M->Type, "Drucker-Prager") == 0) out->type = NLPS_MAT_DRUCKER_PRAGER;
  else if (strcmp(M->Type, "Matsuoka-Nakai") == 0) out->type = NLPS_MAT_MATSUOKA_NAKAI;
  else if (strcmp(M->Type, "Von-Mises") == 0) out->type = NLPS_MAT_VON_MISES;
  else if (strcmp(M->Type, "Hencky") == 0) out->type = NLPS_MAT_HENCKY;
  else if (strcmp(M->Type, "Lade-Duncan") == 0) out->type = NLPS_MAT_LADE_DUNCAN;
  else {
    /* same wording as Constitutive.c:250-254 */
    fprintf(stderr, "%s : %s %s %s \n", "Error in U_Verlet() [B200]", "The material", M->Type,
            "has not been yet implemnented");
    return EXIT_FAILURE;
  }
  out->rho = M->rho;
  out->E = M->E;
  out->nu = M->nu;
  out->reference_pressure = M->ReferencePressure;
  out->kappa_0 = M->kappa_0;
  out->hardening_modulus = M->Hardening_modulus;
  out->plastic_strain_0 = M->Plastic_Strain_0;
  out->phi_frictional = M->phi_Frictional;
  out->psi_frictional = M->psi_Frictional;
  out->exponent_hardening_ortiz = M->Exponent_Hardening_Ortiz;
  out->cohesion = M->Cohesion;
  out->alpha_hardening_borja = M->alpha_Hardening_Borja;
  for (int k = 0; k < 3; k++) out->a_hardening_borja[k] = M->a_Hardening_Borja[k];
  out->theta_hardening_voce = M->theta_Hardening_Voce;
  out->k_0_hardening_voce = M->K_0_Hardening_Voce;
  out->k_inf_hardening_voce = M->K_inf_Hardening_Voce;
  out->delta_hardening_voce = M->delta_Hardening_Voce;
  return EXIT_SUCCESS;
}

/* rebuild Particle.ListNodes with the reference's own allocator so that the driver's
 * free_table__SetLib__ stays valid and list-reading output paths see the device lists */
static void lists_to_chains(Particle MPM_Mesh, const int *counts, const int *lists, int cap) {
  for (int p = 0; p < MPM_Mesh.NumGP; p++) {
    free__SetLib__(&MPM_Mesh.ListNodes[p]);
    MPM_Mesh.ListNodes[p] = NULL;
    for (int k = counts[p] - 1; k >= 0; k--) push__SetLib__(&MPM_Mesh.ListNodes[p], lists[(size_t)p * cap + k]);
    MPM_Mesh.NumberNodes[p] = counts[p];
  }
}


typedef struct b200_inputs {
  nlps_mesh mesh;
  nlps_solver solver;
  nlps_load *bounds, *neumann;
  double *gravity;
  nlps_material *mats;
  nlps_particles st;
  int *r1p, *r1i, *r2p, *r2i;
  int n_bounds, n_neumann;
} b200_inputs;

/* every array the engine needs, from the reference's structs (no copies of the field buffers: Matrix.nV as is) */
static int b200_flatten(b200_inputs *in, Mesh FEM_Mesh, Particle MPM_Mesh, Time_Int_Params Parameters_Solver) {
  const int Ndim = NumberDimensions;
  const int NumTimeStep = Parameters_Solver.NumTimeStep;
  const int Np = MPM_Mesh.NumGP;
  memset(in, 0, sizeof(*in));
  int *r1p, *r1i, *r2p, *r2i;
  nlps_mesh mesh;
  /* ---- mesh */
  chains_to_csr(FEM_Mesh.NodalLocality_0, FEM_Mesh.NumNodesMesh, &r1p, &r1i);
  chains_to_csr(FEM_Mesh.NodalLocality, FEM_Mesh.NumNodesMesh, &r2p, &r2i);
  mesh.ndim = Ndim;
  mesh.n_nodes = FEM_Mesh.NumNodesMesh;
  mesh.coords = FEM_Mesh.Coordinates.nV;
  mesh.ring1_ptr = r1p; mesh.ring1_idx = r1i;
  mesh.ring2_ptr = r2p; mesh.ring2_idx = r2i;
  mesh.h_avg = FEM_Mesh.h_avg;
  mesh.delta_x = FEM_Mesh.DeltaX;

  /* ---- solver parameters + the globals the scheme reads (Globals.h:16-109) */
  nlps_solver solver;
  memset(&solver, 0, sizeof(solver));
  solver.cfl = Parameters_Solver.CFL;
  solver.cel = Parameters_Solver.Cel;
  solver.initial_step = Parameters_Solver.InitialTimeStep;
  solver.num_steps = NumTimeStep;
  solver.gamma_lme = gamma_LME;
  solver.tol_zero_lme = TOL_zero_LME;
  solver.tol_wrapper_lme = TOL_wrapper_LME;
  solver.max_iter_lme = max_iter_LME;
  solver.tol_radial_returning = TOL_Radial_Returning;
  solver.max_iter_radial_returning = Max_Iterations_Radial_Returning;
  solver.thickness = Thickness_Plain_Stress;
  solver.quirk_transposed_eigvec = -1;
  solver.compute_c_ep = 0;
  /* GramsShapeFun (Type=aLME): Particle.Beta is Np x Ndim^2 and Particle.Cut_off_Ellipsoid exists
   * (Generate-One-Phase-Analysis.c:192-202) */
  solver.shape_function = strcmp(ShapeFunctionGP, "aLME") == 0 ? NLPS_SHAPE_ALME : NLPS_SHAPE_LME;

  /* ---- loads */
  nlps_load *bounds = loads_to_pod(&FEM_Mesh.Bounds, NumTimeStep);
  nlps_load *neumann = loads_to_pod(&MPM_Mesh.Neumann_Contours, NumTimeStep);
  double *gravity = NULL;
  if (gravity_field.STATUS == true) { /* U-Newmark-beta.c:1539-1543 */
    gravity = (double *)calloc((size_t)Ndim * NumTimeStep, sizeof(double));
    for (int k = 0; k < Ndim; k++)
      for (int s = 0; s < NumTimeStep; s++) gravity[k * NumTimeStep + s] = gravity_field.Value[k].Fx[s];
  }

  /* ---- materials */
  nlps_material *mats = (nlps_material *)calloc(MPM_Mesh.NumberMaterials, sizeof(nlps_material));
  for (int m = 0; m < MPM_Mesh.NumberMaterials; m++)
    if (material_to_pod(&MPM_Mesh.Mat[m], &mats[m]) == EXIT_FAILURE) {
      free(mats); free(gravity);
      free_loads(bounds, FEM_Mesh.Bounds.NumBounds);
      free_loads(neumann, MPM_Mesh.Neumann_Contours.NumBounds);
      free(r1p); free(r1i); free(r2p); free(r2i);
      return EXIT_FAILURE;
    }

  /* ---- particle fields: the reference's own buffers */
  nlps_particles st;
  memset(&st, 0, sizeof(st));
  Fields *Phi = &MPM_Mesh.Phi;
  st.n = Np;
  st.x_GC = Phi->x_GC.nV; st.dis = Phi->dis.nV; st.D_dis = Phi->D_dis.nV;
  st.vel = Phi->vel.nV; st.acc = Phi->acc.nV;
  st.F_n = Phi->F_n.nV; st.F_n1 = Phi->F_n1.nV; st.DF = Phi->DF.nV;
  st.b_e_n = Phi->b_e_n.nV; st.b_e_n1 = Phi->b_e_n1.nV; st.Stress = Phi->Stress.nV;
  st.C_ep = Phi->C_ep.nV;
  st.J_n = Phi->J_n.nV; st.J_n1 = Phi->J_n1.nV; st.mass = Phi->mass.nV; st.rho = Phi->rho.nV;
  st.Vol_0 = Phi->Vol_0.nV; st.W = Phi->W;
  st.EPS_n = Phi->EPS_n; st.EPS_n1 = Phi->EPS_n1; st.Kappa_n = Phi->Kappa_n; st.Kappa_n1 = Phi->Kappa_n1;
  st.lambda = MPM_Mesh.lambda.nV; st.Beta = MPM_Mesh.Beta.nV;
  st.Cut_off_Ellipsoid = solver.shape_function == NLPS_SHAPE_ALME ? MPM_Mesh.Cut_off_Ellipsoid.nV : NULL;
  st.I0 = MPM_Mesh.I0; st.NumberNodes = MPM_Mesh.NumberNodes; st.MatIdx = MPM_Mesh.MatIdx;
  st.Back_stress = Phi->Back_stress.nV; /* Von-Mises kinematic hardening (U-Analisys.c:152, Constitutive.c:116) */
  /* 3D Neumann loads act on Phi.Area_0 (U-Verlet.c:847-849); the reference declares the field (Types.h:196) and never
   * allocates it, so a 3D deck with loads fails loudly in nlps_b200_create instead of reading a volume as an area */
#if NumberDimensions == 3
  st.Area_0 = Phi->Area_0.nV;
#else
  st.Area_0 = NULL;
#endif

  in->mesh = mesh; in->solver = solver; in->bounds = bounds; in->neumann = neumann; in->gravity = gravity;
  in->mats = mats; in->st = st; in->r1p = r1p; in->r1i = r1i; in->r2p = r2p; in->r2i = r2i;
  in->n_bounds = FEM_Mesh.Bounds.NumBounds; in->n_neumann = MPM_Mesh.Neumann_Contours.NumBounds;
  return EXIT_SUCCESS;
}

/* The nodal results file of a results step (nodal_results_vtk__InOutFun__, InOutFun/Outputs/WriteVtk.c:270-430, called by
 * output_selector U-Verlet.c:1134): mesh, the ActiveNodes mask and the reactions.  The scheme shims step on while the
 * files of step k are written, so the two nodal arrays of step k are captured (two small device -> host copies, n_nodes x
 * (d doubles + 1 byte)) before the next chunk of steps is enqueued and written later, beside the particle file.
 * NLPS_B200_NO_NODAL_VTK=1 skips the file. */
typedef struct b200_nodal {
  int nn, d, valid;
  double *R;
  unsigned char *act;
} b200_nodal;

static int b200_nodal_wanted(void) {
  const char *s = getenv("NLPS_B200_NO_NODAL_VTK");
  return !(s && atoi(s) != 0);
}

static int b200_nodal_capture(nlps_engine *eng, b200_nodal *s, int nn, int d) {
  s->valid = 0;
  if (!b200_nodal_wanted()) return EXIT_SUCCESS;
  if (s->R == NULL) {
    s->nn = nn; s->d = d;
    s->R = (double *)malloc(sizeof(double) * (size_t)nn * d);
    s->act = (unsigned char *)malloc((size_t)nn);
    if (!s->R || !s->act) return EXIT_FAILURE;
  }
  if (nlps_b200_get_nodal(eng, 4, s->R) != EXIT_SUCCESS || nlps_b200_get_active(eng, s->act) != EXIT_SUCCESS)
    return EXIT_FAILURE;
  s->valid = 1;
  return EXIT_SUCCESS;
}

static void b200_nodal_write(b200_nodal *s, Mesh FEM_Mesh, int TimeStep_i, int ResultsTimeStep_) {
  if (!s->valid) return;
  /* generate_NodalMask__MeshTools__ (Nodes-Tools.c:46-84): active nodes numbered in node order, -1 elsewhere */
  Mask ActiveNodes;
  ActiveNodes.Nodes2Mask = (int *)malloc(sizeof(int) * (size_t)s->nn);
  int na = 0;
  for (int i = 0; i < s->nn; i++) ActiveNodes.Nodes2Mask[i] = s->act[i] ? na++ : -1;
  ActiveNodes.Nactivenodes = na;
  Matrix Reactions = allocZ__MatrixLib__(na > 0 ? na : 1, s->d);
  for (int i = 0; i < s->nn; i++) {
    const int m = ActiveNodes.Nodes2Mask[i];
    if (m < 0) continue;
    for (int k = 0; k < s->d; k++) Reactions.nM[m][k] = s->R[(size_t)i * s->d + k];
  }
  nodal_results_vtk__InOutFun__(FEM_Mesh, ActiveNodes, Reactions, TimeStep_i, ResultsTimeStep_);
  free__MatrixLib__(Reactions);
  free(ActiveNodes.Nodes2Mask);
  s->valid = 0;
}

static void b200_nodal_release(b200_nodal *s) {
  free(s->R); free(s->act);
  s->R = NULL; s->act = NULL; s->valid = 0;
}

static void b200_release(b200_inputs *in) {
  free(in->r1p); free(in->r1i); free(in->r2p); free(in->r2i);
  free_loads(in->bounds, in->n_bounds);
  free_loads(in->neumann, in->n_neumann);
  free(in->gravity);
  free(in->mats);
}

/* final state into the caller's structs, as the CPU scheme would leave it */
static int b200_finish(nlps_engine *eng, b200_inputs *in, Mesh FEM_Mesh, Particle MPM_Mesh) {
  int STATUS = EXIT_SUCCESS;
  const int cap = nlps_b200_list_capacity(eng), Np = MPM_Mesh.NumGP;
  int *counts = (int *)malloc(sizeof(int) * Np);
  int *lists = (int *)malloc(sizeof(int) * (size_t)Np * cap);
  if (nlps_b200_download(eng, &in->st) != EXIT_SUCCESS) STATUS = EXIT_FAILURE;
  if (nlps_b200_get_lists(eng, counts, lists, cap) == EXIT_SUCCESS) lists_to_chains(MPM_Mesh, counts, lists, cap);
  unsigned char *act = (unsigned char *)malloc(FEM_Mesh.NumNodesMesh);
  if (nlps_b200_get_active(eng, act) == EXIT_SUCCESS)
    for (int i = 0; i < FEM_Mesh.NumNodesMesh; i++) FEM_Mesh.ActiveNode[i] = act[i] != 0;
  free(act); free(counts); free(lists);
  return STATUS;
}
#endif /* B200_FLATTEN_H */
