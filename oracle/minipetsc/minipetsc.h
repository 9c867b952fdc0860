/*
 * oracle/minipetsc/minipetsc.h -- TEST INFRASTRUCTURE ONLY (never part of the product path).
 *
 * A functional, sequential, dense stand-in for the PETSc objects the reference's implicit schemes use
 * (Formulations/Displacements/U-Newmark-beta.c:160-425, U-Static.c:83-322): Vec, Mat (SeqAIJ), IS, SNES (NEWTONLS),
 * KSP / PC handles.  PETSc itself is a third-party dependency that is ABSENT from /root/reference and from this image
 * (nl-partsol/CMakeLists.txt:91-101 finds it with pkg-config, version unpinned), so the reference's scheme TUs
 * cannot link against it here.  With this shim the reference's OWN U-Newmark-beta.c / U-Static.c compile unmodified
 * from where they lie and run: every stage function of the scheme (lumped mass :528, nodal v_n / a_n :615, Dirichlet
 * list :778, initial guess :879, residual :970, tangent :1646, kinetic increments :1859, particle updates :1917-2072)
 * is the reference's compiled code; only the two library algorithms below are restated:
 *   - SNESSolve (NEWTONLS): Newton with a backtracking line search on |F| (PETSc: cubic backtracking, SNESLINESEARCHBT;
 *     here: step halving, as in oracle/nlps_oracle.c and the CUDA engine), stopped at |F| < abstol or |F| < rtol |F0| or
 *     max_it -- PETSc's stol test (step-length convergence, default 1e-8) is NOT applied, so that the iteration runs to
 *     the residual tolerance: what the goldens pin is the CONVERGED state, which does not depend on the path to it;
 *   - KSPSolve: dense LU with partial pivoting instead of GMRES(30) + Jacobi at rtol 1e-5 (PETSc's defaults, never
 *     overridden by the reference, U-Newmark-beta.c:323-334) -- again: converged states are compared.
 * Matrices are stored dense (n x n): goldens are generated on decks of a few hundred particles.
 */
#ifndef NLPS_MINIPETSC_H
#define NLPS_MINIPETSC_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int PetscErrorCode;
typedef int PetscInt;
typedef double PetscScalar;
typedef double PetscReal;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscBool;
typedef int MPI_Comm;
#define PETSC_COMM_WORLD 0
#define PETSC_COMM_SELF 1
#define PETSC_DECIDE (-1)
#define PETSC_DETERMINE (-1)
#define PETSC_DEFAULT (-2)
#define PETSC_NULL NULL

#define PetscCall(x)                 \
  do {                               \
    PetscErrorCode ierr_ = (x);      \
    if (ierr_) return ierr_;         \
  } while (0)
#define PetscMalloc(bytes, pp) ((*(pp) = malloc((size_t)(bytes) ? (size_t)(bytes) : 1)) ? 0 : 55)
#define PetscMalloc1(n, pp) ((*(pp) = malloc(((size_t)(n) ? (size_t)(n) : 1) * sizeof(**(pp)))) ? 0 : 55)
#define PetscCalloc1(n, pp) ((*(pp) = calloc((size_t)(n) ? (size_t)(n) : 1, sizeof(**(pp)))) ? 0 : 55)
#define PetscFree(p) (free(p), (p) = NULL, 0)
#define PetscPrintf(comm, ...) (printf(__VA_ARGS__), 0)

typedef enum { INSERT_VALUES = 1, ADD_VALUES = 2 } InsertMode;
typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_INFINITY = 3 } NormType;
typedef enum { VEC_IGNORE_OFF_PROC_ENTRIES, VEC_IGNORE_NEGATIVE_INDICES, VEC_SUBSET_OFF_PROC_ENTRIES } VecOption;
typedef enum { MAT_IGNORE_ZERO_ENTRIES = 1, MAT_SYMMETRIC, MAT_NEW_NONZERO_ALLOCATION_ERR, MAT_SPD } MatOption;
typedef enum { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 } MatAssemblyType;
typedef enum { PETSC_COPY_VALUES, PETSC_OWN_POINTER, PETSC_USE_POINTER } PetscCopyMode;

typedef struct mp_vec_ *Vec;
typedef struct mp_mat_ *Mat;
typedef struct mp_is_ *IS;
typedef struct mp_snes_ *SNES;
typedef struct mp_ksp_ *KSP;
typedef struct mp_pc_ *PC;
typedef const char *SNESType;
typedef const char *PCType;
typedef const char *MatSolverType;
#define SNESNEWTONLS "newtonls"
#define PCJACOBI "jacobi"
#define PCCHOLESKY "cholesky"
#define MATSOLVERCHOLMOD "cholmod"

typedef enum {
  SNES_CONVERGED_FNORM_ABS = 2,
  SNES_CONVERGED_FNORM_RELATIVE = 3,
  SNES_CONVERGED_SNORM_RELATIVE = 4,
  SNES_CONVERGED_ITS = 5,
  SNES_DIVERGED_FUNCTION_DOMAIN = -1,
  SNES_DIVERGED_LINEAR_SOLVE = -3,
  SNES_DIVERGED_MAX_IT = -5,
  SNES_DIVERGED_LINE_SEARCH = -6,
  SNES_CONVERGED_ITERATING = 0
} SNESConvergedReason;
extern const char *const *SNESConvergedReasons; /* indexable by a (possibly negative) reason, as PETSc's */

/* Vec */
PetscErrorCode VecCreate(MPI_Comm, Vec *);
PetscErrorCode VecSetSizes(Vec, PetscInt, PetscInt);
PetscErrorCode VecSetFromOptions(Vec);
PetscErrorCode VecSetOption(Vec, VecOption, PetscBool);
PetscErrorCode VecDuplicate(Vec, Vec *);
PetscErrorCode VecDestroy(Vec *);
PetscErrorCode VecSetValues(Vec, PetscInt, const PetscInt *, const PetscScalar *, InsertMode);
PetscErrorCode VecAssemblyBegin(Vec);
PetscErrorCode VecAssemblyEnd(Vec);
PetscErrorCode VecZeroEntries(Vec);
PetscErrorCode VecGetArray(Vec, PetscScalar **);
PetscErrorCode VecRestoreArray(Vec, PetscScalar **);
PetscErrorCode VecGetArrayRead(Vec, const PetscScalar **);
PetscErrorCode VecRestoreArrayRead(Vec, const PetscScalar **);
PetscErrorCode VecPointwiseDivide(Vec w, Vec x, Vec y);
PetscErrorCode VecNorm(Vec, NormType, PetscReal *);
PetscErrorCode VecGetSize(Vec, PetscInt *);
/* IS */
PetscErrorCode ISCreateGeneral(MPI_Comm, PetscInt, const PetscInt *, PetscCopyMode, IS *);
PetscErrorCode ISDestroy(IS *);
/* Mat */
PetscErrorCode MatCreateSeqAIJ(MPI_Comm, PetscInt, PetscInt, PetscInt, const PetscInt *, Mat *);
PetscErrorCode MatCreateAIJ(MPI_Comm, PetscInt, PetscInt, PetscInt, PetscInt, PetscInt, const PetscInt *, PetscInt,
                            const PetscInt *, Mat *);
PetscErrorCode MatSetOption(Mat, MatOption, PetscBool);
PetscErrorCode MatSetFromOptions(Mat);
PetscErrorCode MatZeroEntries(Mat);
PetscErrorCode MatSetValues(Mat, PetscInt, const PetscInt *, PetscInt, const PetscInt *, const PetscScalar *, InsertMode);
PetscErrorCode MatAssemblyBegin(Mat, MatAssemblyType);
PetscErrorCode MatAssemblyEnd(Mat, MatAssemblyType);
PetscErrorCode MatZeroRowsColumnsIS(Mat, IS, PetscScalar, Vec, Vec);
PetscErrorCode MatDestroy(Mat *);
/* SNES / KSP / PC */
typedef PetscErrorCode (*mp_snes_function)(SNES, Vec, Vec, void *);
typedef PetscErrorCode (*mp_snes_jacobian)(SNES, Vec, Mat, Mat, void *);
PetscErrorCode SNESCreate(MPI_Comm, SNES *);
PetscErrorCode SNESSetType(SNES, SNESType);
PetscErrorCode SNESSetOptionsPrefix(SNES, const char *);
PetscErrorCode SNESSetFunction(SNES, Vec, mp_snes_function, void *);
PetscErrorCode SNESSetJacobian(SNES, Mat, Mat, mp_snes_jacobian, void *);
PetscErrorCode SNESGetKSP(SNES, KSP *);
PetscErrorCode SNESSetTolerances(SNES, PetscReal abstol, PetscReal rtol, PetscReal stol, PetscInt maxit, PetscInt maxf);
PetscErrorCode SNESSetLagJacobian(SNES, PetscInt);
PetscErrorCode SNESSetFromOptions(SNES);
PetscErrorCode SNESSolve(SNES, Vec b, Vec x);
PetscErrorCode SNESGetConvergedReason(SNES, SNESConvergedReason *);
PetscErrorCode SNESGetIterationNumber(SNES, PetscInt *);
PetscErrorCode SNESGetLinearSolveIterations(SNES, PetscInt *);
PetscErrorCode SNESDestroy(SNES *);
PetscErrorCode KSPGetPC(KSP, PC *);
PetscErrorCode KSPSetTolerances(KSP, PetscReal, PetscReal, PetscReal, PetscInt);
PetscErrorCode KSPGetResidualNorm(KSP, PetscReal *);
PetscErrorCode PCSetType(PC, PCType);
PetscErrorCode PCFactorSetMatSolverType(PC, MatSolverType);

/* read-outs for the harness (oracle/ref_harness_newmark.c): totals over all SNESSolve calls since the last reset */
void minipetsc_stats(int *solves, int *newton_iters, int *function_evals, int *not_converged, double *last_fnorm);
void minipetsc_reset_stats(void);
#endif
