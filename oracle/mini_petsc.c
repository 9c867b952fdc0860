/*
 * oracle/mini_petsc.c -- TEST INFRASTRUCTURE ONLY (never part of the product path).
 *
 * Implementation of oracle/minipetsc/minipetsc.h: the subset of PETSc that the reference's implicit schemes call
 * (Formulations/Displacements/U-Newmark-beta.c:160-425, U-Static.c:83-322), sequential and dense, so that the reference's
 * own scheme TUs run in this image.  PETSc (absent, version unpinned: nl-partsol/CMakeLists.txt:91-101) semantics that the
 * schemes rely on and that are kept here:
 *   - VecSetValues / MatSetValues with ADD_VALUES; negative indices are skipped (Vec: with VEC_IGNORE_NEGATIVE_INDICES,
 *     U-Newmark-beta.c:280,283,638; Mat: always, as PETSc documents);
 *   - MatZeroRowsColumnsIS(A, is, diag, NULL, NULL): rows and columns zeroed, diag on the diagonal (:1827);
 *   - SNESSolve calls the residual callback with the work vector handed to SNESSetFunction and the Jacobian callback
 *     with the matrix handed to SNESSetJacobian, Jacobian rebuilt at every iteration (SNESSetLagJacobian(1), :343);
 *     the LAST residual evaluation of a solve is at the accepted iterate (the schemes read the particle state it left).
 * Restated library algorithms (see the header): Newton + step halving on |F| for NEWTONLS, dense LU for the KSP.
 */
#include "minipetsc/minipetsc.h"

#include <math.h>

struct mp_vec_ { PetscInt n; PetscScalar *a; int ignore_negative; };
struct mp_mat_ { PetscInt n; PetscScalar *a; };
struct mp_is_ { PetscInt n; const PetscInt *idx; PetscInt *own; };
struct mp_pc_ { int dummy; };
struct mp_ksp_ { struct mp_pc_ pc; PetscReal rnorm; };
struct mp_snes_ {
  Vec r; mp_snes_function fn; void *fctx;
  Mat J; mp_snes_jacobian jac; void *jctx;
  PetscReal abstol, rtol; PetscInt maxit;
  struct mp_ksp_ ksp;
  SNESConvergedReason reason; PetscInt its, lits;
};

static const char *const reasons_[] = {
    "DIVERGED_TR_DELTA", "DIVERGED_JACOBIAN_DOMAIN", "DIVERGED_DTOL", "DIVERGED_LOCAL_MIN", "DIVERGED_INNER",
    "DIVERGED_LINE_SEARCH", "DIVERGED_MAX_IT", "DIVERGED_FNORM_NAN", "DIVERGED_LINEAR_SOLVE", "DIVERGED_FUNCTION_COUNT",
    "DIVERGED_FUNCTION_DOMAIN", "CONVERGED_ITERATING", "CONVERGED_UNUSED", "CONVERGED_FNORM_ABS", "CONVERGED_FNORM_RELATIVE",
    "CONVERGED_SNORM_RELATIVE", "CONVERGED_ITS", "CONVERGED_TR_DELTA"};
const char *const *SNESConvergedReasons = reasons_ + 11;

static struct { int solves, iters, fevals, not_converged; double last_fnorm; } g_stats;
void minipetsc_stats(int *solves, int *newton_iters, int *function_evals, int *not_converged, double *last_fnorm) {
  if (solves) *solves = g_stats.solves;
  if (newton_iters) *newton_iters = g_stats.iters;
  if (function_evals) *function_evals = g_stats.fevals;
  if (not_converged) *not_converged = g_stats.not_converged;
  if (last_fnorm) *last_fnorm = g_stats.last_fnorm;
}
void minipetsc_reset_stats(void) { memset(&g_stats, 0, sizeof(g_stats)); }

/* ------------------------------------------------ Vec */
PetscErrorCode VecCreate(MPI_Comm c, Vec *v) { (void)c; *v = calloc(1, sizeof(**v)); return *v ? 0 : 55; }
PetscErrorCode VecSetSizes(Vec v, PetscInt nloc, PetscInt n) {
  if (n < 0) n = nloc;
  if (n < 0) return 63;
  free(v->a);
  v->n = n;
  v->a = calloc((size_t)(n ? n : 1), sizeof(PetscScalar));
  return v->a ? 0 : 55;
}
PetscErrorCode VecSetFromOptions(Vec v) { (void)v; return 0; }
PetscErrorCode VecSetOption(Vec v, VecOption o, PetscBool f) { if (o == VEC_IGNORE_NEGATIVE_INDICES) v->ignore_negative = (f == PETSC_TRUE); return 0; }
PetscErrorCode VecDuplicate(Vec v, Vec *w) {
  PetscCall(VecCreate(0, w));
  PetscCall(VecSetSizes(*w, PETSC_DECIDE, v->n));
  (*w)->ignore_negative = v->ignore_negative;
  return 0;
}
PetscErrorCode VecDestroy(Vec *v) { if (v && *v) { free((*v)->a); free(*v); *v = NULL; } return 0; }
PetscErrorCode VecSetValues(Vec v, PetscInt ni, const PetscInt *ix, const PetscScalar *y, InsertMode m) {
  for (PetscInt i = 0; i < ni; i++) {
    if (ix[i] < 0) { if (v->ignore_negative) continue; return 63; }
    if (ix[i] >= v->n) return 63;
    if (m == ADD_VALUES) v->a[ix[i]] += y[i]; else v->a[ix[i]] = y[i];
  }
  return 0;
}
PetscErrorCode VecAssemblyBegin(Vec v) { (void)v; return 0; }
PetscErrorCode VecAssemblyEnd(Vec v) { (void)v; return 0; }
PetscErrorCode VecZeroEntries(Vec v) { memset(v->a, 0, sizeof(PetscScalar) * (size_t)v->n); return 0; }
PetscErrorCode VecGetArray(Vec v, PetscScalar **a) { *a = v->a; return 0; }
PetscErrorCode VecRestoreArray(Vec v, PetscScalar **a) { (void)v; if (a) *a = NULL; return 0; }
PetscErrorCode VecGetArrayRead(Vec v, const PetscScalar **a) { *a = v->a; return 0; }
PetscErrorCode VecRestoreArrayRead(Vec v, const PetscScalar **a) { (void)v; if (a) *a = NULL; return 0; }
PetscErrorCode VecPointwiseDivide(Vec w, Vec x, Vec y) {
  if (w->n != x->n || w->n != y->n) return 60;
  for (PetscInt i = 0; i < w->n; i++) w->a[i] = x->a[i] / y->a[i];
  return 0;
}
static double norm2_(const PetscScalar *a, PetscInt n) { double s = 0; for (PetscInt i = 0; i < n; i++) s += a[i] * a[i]; return sqrt(s); }
PetscErrorCode VecNorm(Vec v, NormType t, PetscReal *out) {
  if (t == NORM_2) { *out = norm2_(v->a, v->n); return 0; }
  double s = 0;
  for (PetscInt i = 0; i < v->n; i++) { double f = fabs(v->a[i]); if (t == NORM_1) s += f; else if (f > s) s = f; }
  *out = s;
  return 0;
}
PetscErrorCode VecGetSize(Vec v, PetscInt *n) { *n = v->n; return 0; }

/* ------------------------------------------------ IS */
PetscErrorCode ISCreateGeneral(MPI_Comm c, PetscInt n, const PetscInt *idx, PetscCopyMode mode, IS *is) {
  (void)c;
  *is = calloc(1, sizeof(**is));
  if (!*is) return 55;
  (*is)->n = n;
  if (mode == PETSC_COPY_VALUES) {
    (*is)->own = malloc(sizeof(PetscInt) * (size_t)(n ? n : 1));
    memcpy((*is)->own, idx, sizeof(PetscInt) * (size_t)n);
    (*is)->idx = (*is)->own;
  } else {
    (*is)->idx = idx;
    if (mode == PETSC_OWN_POINTER) (*is)->own = (PetscInt *)idx;
  }
  return 0;
}
PetscErrorCode ISDestroy(IS *is) { if (is && *is) { free((*is)->own); free(*is); *is = NULL; } return 0; }

/* ------------------------------------------------ Mat (dense storage behind the SeqAIJ calls) */
PetscErrorCode MatCreateSeqAIJ(MPI_Comm c, PetscInt m, PetscInt n, PetscInt nz, const PetscInt *nnz, Mat *A) {
  (void)c; (void)nz; (void)nnz;
  if (m != n) return 60;
  *A = calloc(1, sizeof(**A));
  if (!*A) return 55;
  (*A)->n = n;
  (*A)->a = calloc((size_t)n * (size_t)n + 1, sizeof(PetscScalar));
  return (*A)->a ? 0 : 55;
}
PetscErrorCode MatCreateAIJ(MPI_Comm c, PetscInt m, PetscInt n, PetscInt M, PetscInt N, PetscInt dnz, const PetscInt *dnnz,
                            PetscInt onz, const PetscInt *onnz, Mat *A) {
  (void)onz; (void)onnz;
  return MatCreateSeqAIJ(c, M >= 0 ? M : m, N >= 0 ? N : n, dnz, dnnz, A);
}
PetscErrorCode MatSetOption(Mat A, MatOption o, PetscBool f) { (void)A; (void)o; (void)f; return 0; }
PetscErrorCode MatSetFromOptions(Mat A) { (void)A; return 0; }
PetscErrorCode MatZeroEntries(Mat A) { memset(A->a, 0, sizeof(PetscScalar) * (size_t)A->n * (size_t)A->n); return 0; }
PetscErrorCode MatSetValues(Mat A, PetscInt m, const PetscInt *im, PetscInt n, const PetscInt *in, const PetscScalar *v, InsertMode mode) {
  for (PetscInt i = 0; i < m; i++) {
    if (im[i] < 0) continue;
    if (im[i] >= A->n) return 63;
    for (PetscInt j = 0; j < n; j++) {
      if (in[j] < 0) continue;
      if (in[j] >= A->n) return 63;
      PetscScalar *t = &A->a[(size_t)im[i] * A->n + in[j]];
      if (mode == ADD_VALUES) *t += v[i * n + j]; else *t = v[i * n + j];
    }
  }
  return 0;
}
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t) { (void)A; (void)t; return 0; }
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t) { (void)A; (void)t; return 0; }
PetscErrorCode MatZeroRowsColumnsIS(Mat A, IS is, PetscScalar diag, Vec x, Vec b) {
  (void)x; (void)b;
  for (PetscInt k = 0; k < is->n; k++) {
    const PetscInt r = is->idx[k];
    if (r < 0 || r >= A->n) return 63;
    for (PetscInt j = 0; j < A->n; j++) { A->a[(size_t)r * A->n + j] = 0.0; A->a[(size_t)j * A->n + r] = 0.0; }
    A->a[(size_t)r * A->n + r] = diag;
  }
  return 0;
}
PetscErrorCode MatDestroy(Mat *A) { if (A && *A) { free((*A)->a); free(*A); *A = NULL; } return 0; }

/* ------------------------------------------------ KSP / PC handles */
PetscErrorCode KSPGetPC(KSP k, PC *pc) { *pc = &k->pc; return 0; }
PetscErrorCode KSPSetTolerances(KSP k, PetscReal a, PetscReal b, PetscReal c, PetscInt d) { (void)k; (void)a; (void)b; (void)c; (void)d; return 0; }
PetscErrorCode KSPGetResidualNorm(KSP k, PetscReal *r) { *r = k->rnorm; return 0; }
PetscErrorCode PCSetType(PC p, PCType t) { (void)p; (void)t; return 0; }
PetscErrorCode PCFactorSetMatSolverType(PC p, MatSolverType t) { (void)p; (void)t; return 0; }

/* dense LU with partial pivoting; A (n x n, row-major) is destroyed, b becomes the solution */
static int lu_solve_(PetscInt n, PetscScalar *A, PetscScalar *b) {
  for (PetscInt k = 0; k < n; k++) {
    PetscInt piv = k;
    double mx = fabs(A[(size_t)k * n + k]);
    for (PetscInt i = k + 1; i < n; i++) { double f = fabs(A[(size_t)i * n + k]); if (f > mx) { mx = f; piv = i; } }
    if (mx == 0.0) return 1;
    if (piv != k) {
      for (PetscInt j = 0; j < n; j++) { double t = A[(size_t)k * n + j]; A[(size_t)k * n + j] = A[(size_t)piv * n + j]; A[(size_t)piv * n + j] = t; }
      double t = b[k]; b[k] = b[piv]; b[piv] = t;
    }
    const double inv = 1.0 / A[(size_t)k * n + k];
    for (PetscInt i = k + 1; i < n; i++) {
      const double f = A[(size_t)i * n + k] * inv;
      if (f == 0.0) continue;
      for (PetscInt j = k + 1; j < n; j++) A[(size_t)i * n + j] -= f * A[(size_t)k * n + j];
      b[i] -= f * b[k];
    }
  }
  for (PetscInt k = n; k-- > 0;) {
    double s = b[k];
    for (PetscInt j = k + 1; j < n; j++) s -= A[(size_t)k * n + j] * b[j];
    b[k] = s / A[(size_t)k * n + k];
  }
  return 0;
}

/* ------------------------------------------------ SNES */
PetscErrorCode SNESCreate(MPI_Comm c, SNES *s) {
  (void)c;
  *s = calloc(1, sizeof(**s));
  if (!*s) return 55;
  (*s)->abstol = 1e-50; (*s)->rtol = 1e-8; (*s)->maxit = 50; /* PETSc's defaults */
  return 0;
}
PetscErrorCode SNESSetType(SNES s, SNESType t) { (void)s; return strcmp(t, SNESNEWTONLS) ? 56 : 0; }
PetscErrorCode SNESSetOptionsPrefix(SNES s, const char *p) { (void)s; (void)p; return 0; }
PetscErrorCode SNESSetFunction(SNES s, Vec r, mp_snes_function f, void *ctx) { s->r = r; s->fn = f; s->fctx = ctx; return 0; }
PetscErrorCode SNESSetJacobian(SNES s, Mat A, Mat P, mp_snes_jacobian j, void *ctx) { (void)P; s->J = A; s->jac = j; s->jctx = ctx; return 0; }
PetscErrorCode SNESGetKSP(SNES s, KSP *k) { *k = &s->ksp; return 0; }
PetscErrorCode SNESSetTolerances(SNES s, PetscReal abstol, PetscReal rtol, PetscReal stol, PetscInt maxit, PetscInt maxf) {
  (void)stol; (void)maxf;
  if (abstol != (PetscReal)PETSC_DEFAULT) s->abstol = abstol;
  if (rtol != (PetscReal)PETSC_DEFAULT) s->rtol = rtol;
  if (maxit != PETSC_DEFAULT) s->maxit = maxit;
  return 0;
}
PetscErrorCode SNESSetLagJacobian(SNES s, PetscInt lag) { (void)s; return lag == 1 ? 0 : 56; }
PetscErrorCode SNESSetFromOptions(SNES s) { (void)s; return 0; }
PetscErrorCode SNESGetConvergedReason(SNES s, SNESConvergedReason *r) { *r = s->reason; return 0; }
PetscErrorCode SNESGetIterationNumber(SNES s, PetscInt *n) { *n = s->its; return 0; }
PetscErrorCode SNESGetLinearSolveIterations(SNES s, PetscInt *n) { *n = s->lits; return 0; }
PetscErrorCode SNESDestroy(SNES *s) { if (s && *s) { free(*s); *s = NULL; } return 0; }

static int eval_(SNES s, Vec x, Vec f, double *nrm) {
  g_stats.fevals++;
  if (s->fn(s, x, f, s->fctx)) return 1;
  *nrm = norm2_(f->a, f->n);
  return !isfinite(*nrm);
}

PetscErrorCode SNESSolve(SNES s, Vec b, Vec x) {
  if (b) return 56;
  if (!s->fn || !s->jac || !s->r || !s->J || s->J->n != x->n) return 73;
  const PetscInt n = x->n;
  Vec trial, ft;
  PetscCall(VecDuplicate(x, &trial));
  PetscCall(VecDuplicate(x, &ft));
  PetscScalar *delta = malloc(sizeof(PetscScalar) * (size_t)(n ? n : 1));
  PetscScalar *lu = malloc(sizeof(PetscScalar) * ((size_t)n * n + 1));
  PetscErrorCode rc = 0;
  double r0, rn;
  s->its = s->lits = 0;
  s->reason = SNES_CONVERGED_ITERATING;
  g_stats.solves++;
  if (eval_(s, x, s->r, &rn)) { s->reason = SNES_DIVERGED_FUNCTION_DOMAIN; rc = 1; goto done; }
  r0 = rn;
  while (1) {
    if (rn <= s->abstol) { s->reason = SNES_CONVERGED_FNORM_ABS; break; }
    if (rn <= s->rtol * r0) { s->reason = SNES_CONVERGED_FNORM_RELATIVE; break; }
    if (s->its >= s->maxit) { s->reason = SNES_DIVERGED_MAX_IT; break; }
    if (s->jac(s, x, s->J, s->J, s->jctx)) { rc = 1; goto done; }
    memcpy(lu, s->J->a, sizeof(PetscScalar) * (size_t)n * n);
    for (PetscInt k = 0; k < n; k++) delta[k] = -s->r->a[k];
    if (lu_solve_(n, lu, delta)) { s->reason = SNES_DIVERGED_LINEAR_SOLVE; break; }
    s->lits++;
    s->ksp.rnorm = 0.0;
    double lam = 1.0, rt = 0.0;
    int ok = 0;
    for (int ls = 0; ls < 8; ls++, lam *= 0.5) {
      for (PetscInt k = 0; k < n; k++) trial->a[k] = x->a[k] + lam * delta[k];
      if (!eval_(s, trial, ft, &rt) && rt < rn) { ok = 1; break; }
    }
    if (!ok) { /* no decrease: the full step, as a stagnated Newton would take it */
      for (PetscInt k = 0; k < n; k++) trial->a[k] = x->a[k] + delta[k];
      if (eval_(s, trial, ft, &rt)) { s->reason = SNES_DIVERGED_FUNCTION_DOMAIN; rc = 1; goto done; }
    }
    memcpy(x->a, trial->a, sizeof(PetscScalar) * (size_t)n);
    memcpy(s->r->a, ft->a, sizeof(PetscScalar) * (size_t)n);
    s->its++;
    g_stats.iters++;
    const int stalled = !ok && !(rt < rn);
    rn = rt;
    if (stalled) { s->reason = SNES_DIVERGED_LINE_SEARCH; break; }
  }
  g_stats.last_fnorm = rn;
  if (s->reason < 0) g_stats.not_converged++;
done:
  free(delta); free(lu);
  VecDestroy(&trial); VecDestroy(&ft);
  return rc;
}
