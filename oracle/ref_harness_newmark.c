/*
 * oracle/ref_harness_newmark.c -- TEST INFRASTRUCTURE ONLY (never part of the product path).
 *
 * The reference's OWN implicit schemes, run here: Formulations/Displacements/U-Newmark-beta.c (U_Newmark_Beta, :130-432)
 * and U-Static.c (U_Static, :83-322) are compiled unmodified from where they lie against oracle/minipetsc (a functional
 * dense stand-in for the PETSc objects they call; PETSc is absent from /root/reference and from this image) and linked
 * with the other 86 2D translation units into oracle/_ref/libnlps2d_newmark_ref.so (`make -C oracle ref-newmark`).
 * Everything of ref_harness.c (deck parsing through the reference's parser, set-up, field getters) is reused by
 * inclusion, because the reference keeps its simulation in that file's static objects.
 *
 * What this pins: every stage function of the implicit schemes is the reference's compiled code; the Newton loop and
 * the linear solve are oracle/mini_petsc.c (restated library algorithms, see its header).  tests/golden/make_golden.py
 * freezes the converged states (tests/golden/newmark_*.npz); tests/test_implicit_oracle.py checks oracle/nlps_oracle.c's
 * orc_newmark_* restatement against them.
 */
#include "ref_harness.c"

#include "minipetsc/minipetsc.h"

PetscErrorCode U_Newmark_Beta(Mesh, Particle, Time_Int_Params);
PetscErrorCode U_Static(Mesh, Particle, Time_Int_Params);

/* the progress bar of InOutFun/print_ScreenMessage.c:36-63 (that TU is not built: ref_harness.c defines its other helpers) */
void DoProgress(char label[], int step, int total) { (void)label; (void)step; (void)total; }

/* the whole time loop of the deck (InitialTimeStep .. NumTimeStep), as driver-nl-partsol.c:362-375 calls it */
int refh_newmark_run(void) {
  if (!g_ready) return -1;
  minipetsc_reset_stats();
  return (int)U_Newmark_Beta(FEM_Mesh, MPM_Mesh, Params);
}
int refh_static_run(void) {
  if (!g_ready) return -1;
  minipetsc_reset_stats();
  return (int)U_Static(FEM_Mesh, MPM_Mesh, Params);
}
/* out[0..4] = SNES solves, Newton iterations, residual evaluations, solves that did not converge, last |F| */
void refh_newmark_stats(double *out) {
  int a, b, c, d;
  double f;
  minipetsc_stats(&a, &b, &c, &d, &f);
  out[0] = a; out[1] = b; out[2] = c; out[3] = d; out[4] = f;
}
double refh_newmark_tol(void) { return Params.TOL_Newmark_beta; }
int refh_newmark_max_iter(void) { return (int)Params.MaxIter; }
