/*
 * ref_harness3d_lme.c -- TEST INFRASTRUCTURE.  The reference's OWN 3D LME evaluation, callable point-wise.
 *
 * Nodes/LME.c compiles in 3D (NumberDimensions == 3) and so do the helpers it needs (Matlib/MatrixOp.c, MathOp.c,
 * ChainOp.c, Nodes/Nodes-Tools.c); only Matlib/TensorLib.c does not.  This translation unit #includes the
 * reference's LME.c FROM WHERE IT LIES (nothing is copied into the repository), so that its static Newton solver
 * __lambda_Newton_Rapson (LME.c:272-353) is reachable next to the public p__LME__ / dp__LME__ (:700-891), and supplies
 *   - the driver globals LME.c reads (Globals.h:47-51),
 *   - rcond__TensorLib__ (TensorLib.c:965-994: LAPACKE_dlange + LAPACKE_dgecon on the unfactorised matrix), restated here
 *     because its home TU does not compile in 3D; the other TensorLib symbols the helper TUs reference are never reached
 *     from these entry points and abort if they are.
 * Built by `make -C oracle ref3d` into oracle/_ref/libnlps3d_lme_ref.so; pins the 3D LME of the oracle port
 * (tests/golden/lme_points3d.npz, tests/test_oracle_3d_laws.py).
 */
#include "Nodes/LME.c"
#include <lapacke.h> /* oracle/shim: prototypes of the LAPACKE entry points (mini_lapack.c or OpenBLAS) */

#if NumberDimensions != 3
#error "ref_harness3d_lme.c is the 3D build: compile without -DUSE_PLAINSTRAIN"
#endif

int max_iter_LME;
double TOL_zero_LME, TOL_wrapper_LME, gamma_LME;
char wrapper_LME[MAXC];
bool Driver_EigenErosion, Driver_EigenSoftening;
int NumberDOF = 3;

double rcond__TensorLib__(const double *A) {
  double RCOND;
  const double ANORM = LAPACKE_dlange(LAPACK_ROW_MAJOR, '1', 3, 3, A, 3);
  if (LAPACKE_dgecon(LAPACK_ROW_MAJOR, '1', 3, A, 3, ANORM, &RCOND) < 0) return EXIT_FAILURE;
  return RCOND;
}
static void unreachable(const char *what) { fprintf(stderr, "ref_harness3d_lme: %s reached\n", what); abort(); }
Tensor alloc__TensorLib__(int o) { (void)o; unreachable("alloc__TensorLib__"); Tensor t; memset(&t, 0, sizeof(t)); return t; }
Tensor memory_to_tensor__TensorLib__(double *a, int o) { (void)a; (void)o; unreachable("memory_to_tensor__TensorLib__"); Tensor t; memset(&t, 0, sizeof(t)); return t; }
void free__TensorLib__(Tensor a) { (void)a; unreachable("free__TensorLib__"); }
Tensor dyadic_Product__TensorLib__(Tensor a, Tensor b) { (void)b; unreachable("dyadic_Product__TensorLib__"); return a; }

/* n neighbours with l_a = x_p - x_a (n x 3, row-major); lambda: in = start value, out = converged;
 * N[n], dN[n x 3].  Returns the status of the reference's Newton. */
int refh3_lme_point(int n, const double *l, double *lambda, double beta, double tol_wrapper, int max_iter, double *N,
                    double *dN) {
  TOL_wrapper_LME = tol_wrapper;
  max_iter_LME = max_iter;
  strcpy(wrapper_LME, "Newton-Raphson");
  double *lc = (double *)malloc(sizeof(double) * 3 * (size_t)n);
  memcpy(lc, l, sizeof(double) * 3 * (size_t)n);
  Matrix L = memory_to_matrix__MatrixLib__(n, 3, lc);
  Matrix lam = memory_to_matrix__MatrixLib__(3, 1, lambda);
  const int status = __lambda_Newton_Rapson(0, L, lam, beta);
  Matrix p = p__LME__(L, lam, beta);
  Matrix dp = dp__LME__(L, p);
  for (int a = 0; a < n; a++) {
    N[a] = p.nV[a];
    for (int i = 0; i < 3; i++) dN[a * 3 + i] = dp.nM[a][i];
  }
  free__MatrixLib__(p);
  free__MatrixLib__(dp);
  free(L.nM); /* memory_to_matrix allocates a row table for true matrices only (MatrixOp.c:197-216) */
  free(lc);
  return status;
}
double refh3_beta(double gamma, double h_avg) { return beta__LME__(gamma, h_avg); }

/* tributary__LME__ (LME.c:1019-1099, static: reachable because this TU includes LME.c) on one particle: the candidate nodes
 * are the 2-ring of its closest node in chain order, with their coordinates and ActiveNode flags; out = the accepted
 * nodes as LOCAL candidate indices in the order of the returned chain (reverse acceptance order); returns their number. */
int refh3_tributary(int n_cand, const double *coords, const unsigned char *active, const double *xp, double beta,
                    double tol_zero, int *out) {
  TOL_zero_LME = tol_zero;
  Mesh M;
  memset(&M, 0, sizeof(M));
  double *cc = (double *)malloc(sizeof(double) * 3 * (size_t)n_cand);
  memcpy(cc, coords, sizeof(double) * 3 * (size_t)n_cand);
  M.NumNodesMesh = n_cand;
  M.Coordinates = memory_to_matrix__MatrixLib__(n_cand, 3, cc);
  M.ActiveNode = (bool *)malloc(sizeof(bool) * (size_t)n_cand);
  for (int i = 0; i < n_cand; i++) M.ActiveNode[i] = active[i] != 0;
  ChainPtr ring = NULL;
  for (int i = n_cand - 1; i >= 0; i--) push__SetLib__(&ring, i); /* push adds at the head: traversal = 0 .. n_cand-1 */
  ChainPtr locality[1] = {ring};
  int size[1] = {n_cand};
  M.NodalLocality = locality;
  M.SizeNodalLocality = size;
  double x[3] = {xp[0], xp[1], xp[2]};
  Matrix X_p = memory_to_matrix__MatrixLib__(3, 1, x);
  ChainPtr lst = tributary__LME__(0, X_p, beta, 0, M);
  const int n = lenght__SetLib__(lst);
  int *arr = set_to_memory__SetLib__(lst, n);
  memcpy(out, arr, sizeof(int) * (size_t)n);
  free(arr);
  free__SetLib__(&lst);
  free__SetLib__(&ring);
  free(M.Coordinates.nM);
  free(M.ActiveNode);
  free(cc);
  return n;
}

/* get_closest_node__MeshTools__ (Nodes/Nodes-Tools.c:476-538) among the given candidates (the 1-ring of the previous
 * closest node in chain order): returns the LOCAL index of the winner. */
int refh3_closest(int n_cand, const double *coords, const double *xp) {
  double *cc = (double *)malloc(sizeof(double) * 3 * (size_t)n_cand);
  memcpy(cc, coords, sizeof(double) * 3 * (size_t)n_cand);
  Matrix C = memory_to_matrix__MatrixLib__(n_cand, 3, cc);
  ChainPtr ring = NULL;
  for (int i = n_cand - 1; i >= 0; i--) push__SetLib__(&ring, i);
  double x[3] = {xp[0], xp[1], xp[2]};
  Matrix X_p = memory_to_matrix__MatrixLib__(3, 1, x);
  const int best = get_closest_node__MeshTools__(X_p, ring, C);
  free__SetLib__(&ring);
  free(C.nM);
  free(cc);
  return best;
}
