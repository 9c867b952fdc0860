"""oracle/mini_petsc.c -- the dense stand-in for the PETSc calls the reference's implicit schemes make (test
infrastructure: it lets the reference's own U-Newmark-beta.c / U-Static.c run here) -- checked on its own: the Vec / Mat
semantics those files rely on and the Newton solve on a small system with a known root.  CPU only (gcc)."""
import os
import subprocess
import tempfile

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")

PROGRAM = r'''
#include <math.h>
#include <stdio.h>
#include "minipetsc/minipetsc.h"

#define CHECK(c) do { if (!(c)) { printf("FAILED line %d: %s\n", __LINE__, #c); return 1; } } while (0)

/* F(x) = (x0^2 + x1 - 3, x0 + x1^2 - 5), root (1, 2); ctx counts the evaluations */
static PetscErrorCode fn(SNES s, Vec x, Vec f, void *ctx) {
  const PetscScalar *a; PetscScalar *r;
  (void)s; ++*(int *)ctx;
  PetscCall(VecGetArrayRead(x, &a)); PetscCall(VecGetArray(f, &r));
  r[0] = a[0] * a[0] + a[1] - 3.0; r[1] = a[0] + a[1] * a[1] - 5.0;
  PetscCall(VecRestoreArrayRead(x, &a)); PetscCall(VecRestoreArray(f, &r));
  return 0;
}
static PetscErrorCode jac(SNES s, Vec x, Mat J, Mat B, void *ctx) {
  const PetscScalar *a; (void)s; (void)B; (void)ctx;
  PetscCall(VecGetArrayRead(x, &a));
  PetscInt idx[2] = {0, 1};
  PetscScalar v[4] = {2 * a[0], 1.0, 1.0, 2 * a[1]};
  PetscCall(MatZeroEntries(J));
  PetscCall(MatSetValues(J, 2, idx, 2, idx, v, ADD_VALUES));
  PetscCall(MatAssemblyBegin(J, MAT_FINAL_ASSEMBLY)); PetscCall(MatAssemblyEnd(J, MAT_FINAL_ASSEMBLY));
  return 0;
}

/* a linear system whose matrix is only right if MatSetValues skips negative rows / columns, ADD_VALUES accumulates and
 * MatZeroRowsColumnsIS zeroes row and column and leaves the diagonal: A = diag(2, 4, 1), F(x) = A x - (2, 8, 5) */
static PetscErrorCode fn3(SNES s, Vec x, Vec f, void *ctx) {
  const PetscScalar *a; PetscScalar *r; (void)s; (void)ctx;
  PetscCall(VecGetArrayRead(x, &a)); PetscCall(VecGetArray(f, &r));
  r[0] = 2 * a[0] - 2; r[1] = 4 * a[1] - 8; r[2] = a[2] - 5;
  return 0;
}
static PetscErrorCode jac3(SNES s, Vec x, Mat A, Mat B, void *ctx) {
  (void)s; (void)x; (void)B;
  PetscInt r3[3] = {0, -1, 2}; PetscScalar m9[9] = {1, 2, 3, 4, 5, 6, 7, 8, 9};
  PetscCall(MatZeroEntries(A));
  PetscCall(MatSetValues(A, 3, r3, 3, r3, m9, ADD_VALUES));   /* (0,0) 1, (0,2) 3, (2,0) 7, (2,2) 9; row / column -1 skipped */
  PetscCall(MatSetValues(A, 3, r3, 3, r3, m9, ADD_VALUES));   /* twice: 2, 6, 14, 18 */
  PetscInt mid = 1; PetscScalar d = 4.0;
  PetscCall(MatSetValues(A, 1, &mid, 1, &mid, &d, ADD_VALUES));
  PetscCall(MatZeroRowsColumnsIS(A, *(IS *)ctx, 1.0, NULL, NULL)); /* row and column 2 -> unit */
  return 0;
}

int main(void) {
  /* Vec: ADD_VALUES accumulates, negative indices are skipped only with VEC_IGNORE_NEGATIVE_INDICES */
  Vec v, w;
  PetscCall(VecCreate(PETSC_COMM_WORLD, &v)); PetscCall(VecSetSizes(v, PETSC_DECIDE, 4)); PetscCall(VecSetFromOptions(v));
  PetscInt ix[3] = {1, -1, 1}; PetscScalar y[3] = {2.0, 100.0, 0.5};
  CHECK(VecSetValues(v, 3, ix, y, ADD_VALUES) != 0);                 /* PETSc refuses the negative index */
  PetscCall(VecZeroEntries(v));
  PetscCall(VecSetOption(v, VEC_IGNORE_NEGATIVE_INDICES, PETSC_TRUE));
  PetscCall(VecSetValues(v, 3, ix, y, ADD_VALUES));
  const PetscScalar *a; PetscCall(VecGetArrayRead(v, &a));
  CHECK(a[0] == 0.0 && a[1] == 2.5 && a[2] == 0.0 && a[3] == 0.0);
  PetscCall(VecDuplicate(v, &w));
  PetscInt all[4] = {0, 1, 2, 3}; PetscScalar two[4] = {2.0, 2.0, 2.0, 2.0};
  PetscCall(VecSetValues(w, 4, all, two, INSERT_VALUES));
  PetscCall(VecPointwiseDivide(v, v, w));
  PetscReal nrm; PetscCall(VecNorm(v, NORM_2, &nrm));
  CHECK(nrm == 1.25);
  PetscCall(VecDestroy(&v)); PetscCall(VecDestroy(&w));
  /* Mat semantics, observed through a solve (the entries are private): one exact Newton step */
  {
    Mat A; Vec x3, r3v; SNES s3; IS is; PetscInt two_ = 2, it3;
    PetscCall(MatCreateSeqAIJ(PETSC_COMM_SELF, 3, 3, 0, NULL, &A));
    PetscCall(ISCreateGeneral(PETSC_COMM_WORLD, 1, &two_, PETSC_COPY_VALUES, &is));
    PetscCall(VecCreate(PETSC_COMM_WORLD, &x3)); PetscCall(VecSetSizes(x3, PETSC_DECIDE, 3)); PetscCall(VecDuplicate(x3, &r3v));
    PetscCall(SNESCreate(PETSC_COMM_WORLD, &s3));
    PetscCall(SNESSetFunction(s3, r3v, fn3, NULL)); PetscCall(SNESSetJacobian(s3, A, A, jac3, &is));
    PetscCall(SNESSetTolerances(s3, 1e-13, 1e-13, PETSC_DEFAULT, 5, PETSC_DEFAULT));
    PetscCall(SNESSolve(s3, PETSC_NULL, x3));
    PetscCall(SNESGetIterationNumber(s3, &it3));
    PetscCall(VecGetArrayRead(x3, &a));
    CHECK(it3 == 1 && a[0] == 1.0 && a[1] == 2.0 && a[2] == 5.0);
    PetscCall(SNESDestroy(&s3)); PetscCall(ISDestroy(&is)); PetscCall(MatDestroy(&A)); PetscCall(VecDestroy(&x3)); PetscCall(VecDestroy(&r3v));
  }

  /* SNES: Newton with step halving from (3, 3); rtol 1e-12 */
  SNES snes; KSP ksp; PC pc; Vec x, res; Mat J; int evals = 0;
  PetscCall(SNESCreate(PETSC_COMM_WORLD, &snes)); PetscCall(SNESSetType(snes, SNESNEWTONLS));
  PetscCall(VecCreate(PETSC_COMM_WORLD, &x)); PetscCall(VecSetSizes(x, PETSC_DECIDE, 2)); PetscCall(VecDuplicate(x, &res));
  PetscCall(MatCreateSeqAIJ(PETSC_COMM_SELF, 2, 2, 0, NULL, &J));
  PetscCall(SNESSetFunction(snes, res, fn, &evals)); PetscCall(SNESSetJacobian(snes, J, J, jac, NULL));
  PetscCall(SNESGetKSP(snes, &ksp)); PetscCall(KSPGetPC(ksp, &pc)); PetscCall(PCSetType(pc, PCJACOBI));
  PetscCall(SNESSetTolerances(snes, 1e-14, 1e-12, PETSC_DEFAULT, 50, PETSC_DEFAULT));
  PetscCall(SNESSetLagJacobian(snes, 1)); PetscCall(SNESSetFromOptions(snes));
  PetscInt i2[2] = {0, 1}; PetscScalar x0[2] = {3.0, 3.0};
  PetscCall(VecSetValues(x, 2, i2, x0, INSERT_VALUES));
  PetscCall(SNESSolve(snes, PETSC_NULL, x));
  SNESConvergedReason reason; PetscInt its, lits;
  PetscCall(SNESGetConvergedReason(snes, &reason)); PetscCall(SNESGetIterationNumber(snes, &its));
  PetscCall(SNESGetLinearSolveIterations(snes, &lits));
  PetscCall(VecGetArrayRead(x, &a));
  CHECK(reason > 0 && its >= 3 && its <= 10 && lits == its && evals >= its + 1);
  CHECK(fabs(a[0] - 1.0) < 1e-10 && fabs(a[1] - 2.0) < 1e-10);
  PetscCall(VecNorm(res, NORM_2, &nrm));                              /* the work vector holds F at the accepted iterate */
  CHECK(nrm < 1e-10);
  printf("%s its %d evals %d\n", SNESConvergedReasons[reason], its, evals);
  /* a diverged solve is reported, not hidden: max_it = 1 */
  PetscCall(VecSetValues(x, 2, i2, x0, INSERT_VALUES));
  PetscCall(SNESSetTolerances(snes, 1e-14, 1e-12, PETSC_DEFAULT, 1, PETSC_DEFAULT));
  PetscCall(SNESSolve(snes, PETSC_NULL, x));
  PetscCall(SNESGetConvergedReason(snes, &reason));
  CHECK(reason == SNES_DIVERGED_MAX_IT);
  int solves, iters, fe, nc; double last;
  minipetsc_stats(&solves, &iters, &fe, &nc, &last);
  CHECK(solves == 3 && nc == 1 && fe == evals + 2);   /* (+ the linear solve above: two evaluations) */
  PetscCall(SNESDestroy(&snes)); PetscCall(MatDestroy(&J)); PetscCall(VecDestroy(&x)); PetscCall(VecDestroy(&res));
  printf("OK\n");
  return 0;
}
'''


def test_mini_petsc_semantics_and_newton():
    with tempfile.TemporaryDirectory() as tmp:
        src, exe = os.path.join(tmp, "t.c"), os.path.join(tmp, "t")
        with open(src, "w") as f:
            f.write(PROGRAM)
        subprocess.run(["gcc", "-std=gnu11", "-O1", "-Wall", "-I", os.path.join(ROOT, "oracle"), src,
                        os.path.join(ROOT, "oracle", "mini_petsc.c"), "-o", exe, "-lm"], check=True)
        out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
    assert "CONVERGED_FNORM" in out.stdout
