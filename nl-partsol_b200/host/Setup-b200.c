/*
 * Setup-b200.c -- scalable replacements for the two QUADRATIC set-up routines of the reference, so
 * that the reference's own driver reaches the particle counts the B200 engine is built for:
 *
 *   get_sourrounding_elements   InOutFun/Read_GramsBox.c:293-330   O(Nn * Ne) double loop with a malloc per element
 *   initialize__LME__           Nodes/LME.c:45-173                 every particle scans all elements for the one that holds it
 *
 * Both are replaced at LINK time, the reference's sources stay where they lie and untouched: Read_GramsBox.c is compiled
 * with -Dstatic= (its file-local helpers become ordinary symbols), objcopy --weaken-symbol turns the two definitions
 * into weak ones, and the strong definitions below win (nl-partsol_b200/host/Makefile).  What a maintainer would do
 * instead is delete the two function bodies and call these.
 *
 * Results are identical to the reference's: the same chains in the same order (push-front discovery order), the same
 * closest node (get_closest_node__MeshTools__ of the reference itself is called), the first element in index order
 * that holds the particle.  The first neighbour lists, beta and the Lagrange multipliers come from the B200 engine
 * (nlps_b200_initialize_lme == initialize__LME__ phases 2-3), not from the CPU loop.
 *
 * Measured on this container's host (8 threads): reference set-up of a 102,400-particle deck 69 s; with this file: see
 * DESIGN.md section 5.
 */
#include "b200_flatten.h"

/* NodeNeighbour[i] = elements that hold node i, descending element index (the reference pushes ascending j at the head). */
void get_sourrounding_elements(Mesh FEM_Mesh) {
  for (int j = 0; j < FEM_Mesh.NumElemMesh; j++)
    for (ChainPtr c = FEM_Mesh.Connectivity[j]; c != NULL; c = c->next) {
      push__SetLib__(&FEM_Mesh.NodeNeighbour[c->Idx], j);
      FEM_Mesh.NumNeighbour[c->Idx] += 1;
    }
  for (int i = 0; i < FEM_Mesh.NumNodesMesh; i++)
    if (FEM_Mesh.NumNeighbour[i] == 0) {
      fprintf(stderr, "%s : %i \n", "Error computing the sourrounding elements of", i);
      exit(EXIT_FAILURE);
    }
}

/* uniform bucket grid over the bounding boxes of the elements: bucket -> elements in ascending index order */
typedef struct bucket_grid {
  int n[3];
  double lo[3], inv[3];
  int *ptr, *idx;
} bucket_grid;

static void bucket_range(const bucket_grid *g, int d, double a, double b, int *i0, int *i1) {
  int x0 = (int)floor((a - g->lo[d]) * g->inv[d]), x1 = (int)floor((b - g->lo[d]) * g->inv[d]);
  if (x0 < 0) x0 = 0;
  if (x1 > g->n[d] - 1) x1 = g->n[d] - 1;
  *i0 = x0;
  *i1 = x1;
}

static void bucket_build(bucket_grid *g, Mesh FEM_Mesh) {
  const int Ndim = NumberDimensions, Nelem = FEM_Mesh.NumElemMesh, Nn = FEM_Mesh.NumNodesMesh;
  double hi[3] = {0, 0, 0}, ext = 0.0;
  for (int d = 0; d < 3; d++) { g->lo[d] = 0.0; g->n[d] = 1; g->inv[d] = 1.0; }
  for (int d = 0; d < Ndim; d++) {
    g->lo[d] = hi[d] = FEM_Mesh.Coordinates.nM[0][d];
    for (int i = 1; i < Nn; i++) {
      const double x = FEM_Mesh.Coordinates.nM[i][d];
      if (x < g->lo[d]) g->lo[d] = x;
      if (x > hi[d]) hi[d] = x;
    }
  }
  double *elo = (double *)malloc(sizeof(double) * 3 * (size_t)Nelem), *ehi = (double *)malloc(sizeof(double) * 3 * (size_t)Nelem);
  for (int e = 0; e < Nelem; e++)
    for (int d = 0; d < Ndim; d++) {
      double a = 1e300, b = -1e300;
      for (ChainPtr c = FEM_Mesh.Connectivity[e]; c != NULL; c = c->next) {
        const double x = FEM_Mesh.Coordinates.nM[c->Idx][d];
        if (x < a) a = x;
        if (x > b) b = x;
      }
      elo[3 * (size_t)e + d] = a;
      ehi[3 * (size_t)e + d] = b;
      if (b - a > ext) ext = b - a;
    }
  long long nb = 1;
  for (int d = 0; d < Ndim; d++) {  /* buckets about two elements wide */
    g->n[d] = (int)floor((hi[d] - g->lo[d]) / (2.0 * ext)) + 1;
    g->inv[d] = 1.0 / (2.0 * ext);
    nb *= g->n[d];
  }
  g->ptr = (int *)calloc((size_t)nb + 1, sizeof(int));
  for (int pass = 0; pass < 2; pass++) {
    int *fill = pass ? (int *)malloc(sizeof(int) * (size_t)nb) : NULL;
    if (pass) {
      for (long long b = 0; b < nb; b++) g->ptr[b + 1] += g->ptr[b];
      g->idx = (int *)malloc(sizeof(int) * (size_t)(g->ptr[nb] > 0 ? g->ptr[nb] : 1));
      for (long long b = 0; b < nb; b++) fill[b] = g->ptr[b];
    }
    for (int e = 0; e < Nelem; e++) {  /* ascending e: every bucket list ends up in ascending element order */
      int r0[3] = {0, 0, 0}, r1[3] = {0, 0, 0};
      for (int d = 0; d < Ndim; d++) bucket_range(g, d, elo[3 * (size_t)e + d], ehi[3 * (size_t)e + d], &r0[d], &r1[d]);
      for (int k = r0[2]; k <= r1[2]; k++)
        for (int j = r0[1]; j <= r1[1]; j++)
          for (int i = r0[0]; i <= r1[0]; i++) {
            const long long b = ((long long)k * g->n[1] + j) * g->n[0] + i;
            if (pass) g->idx[fill[b]++] = e; else g->ptr[b + 1]++;
          }
    }
    free(fill);
  }
  free(elo);
  free(ehi);
}

/*
 * initialize__LME__ (Nodes/LME.c:45-173).  Phase 1 (:59-118): element that holds the particle (the first in index
 * order, as the reference's scan finds it) and its closest node; phase 2 (:126-145): ActiveNode over the 1-rings;
 * phase 3 (:147-172): first neighbour lists with the beta the particle carries (0 after allocation => every active node
 * of the 2-ring), beta, Newton for lambda -- on the B200.
 */
void initialize__LME__(Particle MPM_Mesh, Mesh FEM_Mesh) {
  const int Ndim = NumberDimensions, Np = MPM_Mesh.NumGP;
  if (strcmp(wrapper_LME, "Nelder-Mead") == 0) {
    fprintf(stderr, "" RED "Error in initialize__LME__() [B200]: only wrapper=Newton-Raphson is supported" RESET " \n");
    exit(EXIT_FAILURE);
  }
  bucket_grid g;
  bucket_build(&g, FEM_Mesh);
  int failed = -1;
#pragma omp parallel for schedule(static)
  for (int p = 0; p < Np; p++) {
    Matrix X_p = memory_to_matrix__MatrixLib__(Ndim, 1, MPM_Mesh.Phi.x_GC.nM[p]);
    int b3[3] = {0, 0, 0}, dummy;
    for (int d = 0; d < Ndim; d++) bucket_range(&g, d, X_p.nV[d], X_p.nV[d], &b3[d], &dummy);
    const long long b = ((long long)b3[2] * g.n[1] + b3[1]) * g.n[0] + b3[0];
    bool found = false;
    for (int q = g.ptr[b]; q < g.ptr[b + 1] && !found; q++) {
      const int e = g.idx[q];
      ChainPtr conn = FEM_Mesh.Connectivity[e];
      Matrix Xe = get_nodes_coordinates__MeshTools__(conn, FEM_Mesh.Coordinates);
      if (FEM_Mesh.In_Out_Element(X_p, Xe) == true) {
        found = true;
        MPM_Mesh.Element_p[p] = e;
        MPM_Mesh.I0[p] = get_closest_node__MeshTools__(X_p, conn, FEM_Mesh.Coordinates);
      }
      free__MatrixLib__(Xe);
    }
    if (!found) {
#pragma omp critical
      failed = p;
    }
  }
  free(g.ptr);
  free(g.idx);
  if (failed >= 0) {
    fprintf(stderr, "%s : %s %i\n", "Error in initialize__LME__()", "The search algorithm was unable to find particle", failed);
    exit(EXIT_FAILURE);
  }
  /* phase 2 (GramsBox leaves ActiveNode uninitialised: all false first, DESIGN.md section 6, deviation 2) */
  for (int i = 0; i < FEM_Mesh.NumNodesMesh; i++) FEM_Mesh.ActiveNode[i] = false;
  for (int p = 0; p < Np; p++) {
    const int I0 = MPM_Mesh.I0[p];
    if ((Driver_EigenErosion == true) || (Driver_EigenSoftening == true)) push__SetLib__(&FEM_Mesh.List_Particles_Node[I0], p);
    for (ChainPtr c = FEM_Mesh.NodalLocality_0[I0]; c != NULL; c = c->next) FEM_Mesh.ActiveNode[c->Idx] = true;
  }
  /* phase 3 on the device: an engine that only knows the mesh, the LME parameters and the particle positions */
  nlps_mesh mesh;
  int *r1p, *r1i, *r2p, *r2i;
  chains_to_csr(FEM_Mesh.NodalLocality_0, FEM_Mesh.NumNodesMesh, &r1p, &r1i);
  chains_to_csr(FEM_Mesh.NodalLocality, FEM_Mesh.NumNodesMesh, &r2p, &r2i);
  mesh.ndim = Ndim; mesh.n_nodes = FEM_Mesh.NumNodesMesh; mesh.coords = FEM_Mesh.Coordinates.nV;
  mesh.ring1_ptr = r1p; mesh.ring1_idx = r1i; mesh.ring2_ptr = r2p; mesh.ring2_idx = r2i;
  mesh.h_avg = FEM_Mesh.h_avg; mesh.delta_x = FEM_Mesh.DeltaX;
  nlps_solver solver;
  memset(&solver, 0, sizeof(solver));
  solver.cfl = 1.0; solver.cel = 1.0; solver.initial_step = 0; solver.num_steps = 1;
  solver.gamma_lme = gamma_LME; solver.tol_zero_lme = TOL_zero_LME; solver.tol_wrapper_lme = TOL_wrapper_LME;
  solver.max_iter_lme = max_iter_LME; solver.thickness = 1.0; solver.quirk_transposed_eigvec = -1;
  const int nmat = MPM_Mesh.NumberMaterials > 0 ? MPM_Mesh.NumberMaterials : 1;
  nlps_material *mats = (nlps_material *)calloc(nmat, sizeof(nlps_material)); /* the laws play no role here */
  for (int m = 0; m < nmat; m++) { mats[m].type = NLPS_MAT_NEO_HOOKEAN_WRIGGERS; mats[m].rho = 1.0; mats[m].E = 1.0; mats[m].nu = 0.25; }
  nlps_particles st;
  memset(&st, 0, sizeof(st));
  st.n = Np;
  st.x_GC = MPM_Mesh.Phi.x_GC.nV; st.mass = MPM_Mesh.Phi.mass.nV; st.rho = MPM_Mesh.Phi.rho.nV; st.Vol_0 = MPM_Mesh.Phi.Vol_0.nV;
  st.lambda = MPM_Mesh.lambda.nV; st.Beta = MPM_Mesh.Beta.nV;
  st.I0 = MPM_Mesh.I0; st.NumberNodes = MPM_Mesh.NumberNodes; st.MatIdx = MPM_Mesh.MatIdx;
  char msg[256];
  nlps_engine *eng = nlps_b200_create(&mesh, &solver, 0, NULL, 0, NULL, NULL, nmat, mats, &st, 0, msg, sizeof(msg));
  if (eng == NULL || nlps_b200_initialize_lme(eng) != EXIT_SUCCESS) {
    fprintf(stderr, "" RED "Error in initialize__LME__() [B200]: %s" RESET " \n", eng == NULL ? msg : "nlps_b200_initialize_lme");
    exit(EXIT_FAILURE);
  }
  nlps_particles out;
  memset(&out, 0, sizeof(out));
  out.n = Np;
  out.lambda = MPM_Mesh.lambda.nV; out.Beta = MPM_Mesh.Beta.nV; out.NumberNodes = MPM_Mesh.NumberNodes;
  const int cap = nlps_b200_list_capacity(eng);
  int *counts = (int *)malloc(sizeof(int) * (size_t)Np), *lists = (int *)malloc(sizeof(int) * (size_t)Np * cap);
  if (nlps_b200_download(eng, &out) != EXIT_SUCCESS || nlps_b200_get_lists(eng, counts, lists, cap) != EXIT_SUCCESS) {
    fprintf(stderr, "" RED "Error in initialize__LME__() [B200]: download" RESET " \n");
    exit(EXIT_FAILURE);
  }
  lists_to_chains(MPM_Mesh, counts, lists, cap);
  free(counts); free(lists);
  nlps_b200_destroy(eng);
  free(mats); free(r1p); free(r1i); free(r2p); free(r2i);
}
