/* Stand-in for <petscksp.h>: PETSc is absent from this image.  It lets the reference's driver
 * compile with -DUSE_PETSC so that its scheme dispatch reaches U_Newmark_Beta / U_Static
 * (driver-nl-partsol.c:254,360-365); no PETSc function is implemented -- the B200 shim
 * (nl-partsol_b200/host/U-Newmark-beta-b200.c) needs none.  Not product code. */
#ifndef NLPS_PETSC_STANDIN_H
#define NLPS_PETSC_STANDIN_H
typedef int PetscErrorCode;
typedef int PetscInt;
typedef double PetscScalar;
typedef void *Mat, *Vec, *KSP, *PC, *SNES, *IS;
static inline PetscErrorCode PetscInitialize(int *argc, char ***argv, const char *file, const char *help) {
  (void)argc; (void)argv; (void)file; (void)help;
  return 0;
}
static inline PetscErrorCode PetscFinalize(void) { return 0; }
#endif
