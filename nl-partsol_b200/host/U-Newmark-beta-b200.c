/*
 * U-Newmark-beta-b200.c -- drop-in replacement for the reference's implicit scheme function.
 *
 * Same symbol, same by-value struct arguments and same return convention as
 *     PetscErrorCode U_Newmark_Beta(Mesh FEM_Mesh, Particle MPM_Mesh, Time_Int_Params Parameters_Solver)
 * (Formulations/Displacements/U-Newmark-beta.c:130, dispatched from driver-nl-partsol.c:360-365 when the
 * deck says `NLPS-Solver (Type=Newmark-beta-Finite-Strains)`; PetscErrorCode is an int).  Linked INSTEAD OF
 * U-Newmark-beta.c: the PETSc objects that file creates every step (Vec / Mat / IS / SNES / KSP / PC,
 * :220-425) live on the B200 behind include/nlps_b200.h (nlps_b200_newmark_*), so this build needs no PETSc
 * library; the driver is compiled with -DUSE_PETSC against the stand-in headers of oracle/shim/ only to
 * switch its dispatch on.  Host glue only, as U-Verlet-b200.c.
 */
#include "b200_flatten.h"
#include "b200_vtk_binary.h" /* NLPS_B200_VTK_BINARY=1: binary twin of the reference's VTK writer */

extern double DeltaTimeStep; /* U-Verlet-b200.c */

static int b200_implicit_scheme(Mesh FEM_Mesh, Particle MPM_Mesh, Time_Int_Params Parameters_Solver, int quasi_static) {
  const int NumTimeStep = Parameters_Solver.NumTimeStep;
  int STATUS = EXIT_SUCCESS;

  if (strcmp(ShapeFunctionGP, "LME") != 0 && (strcmp(ShapeFunctionGP, "aLME") != 0 || NumberDimensions != 2)) {
    fprintf(stderr, "" RED "Error in %s() [B200]: only GramsShapeFun (Type=LME) and, in 2D, (Type=aLME) are supported" RESET " \n",
            quasi_static ? "U_Static" : "U_Newmark_Beta");
    return EXIT_FAILURE;
  }
  b200_inputs in;
  if (b200_flatten(&in, FEM_Mesh, MPM_Mesh, Parameters_Solver) != EXIT_SUCCESS) return EXIT_FAILURE;

  char msg[256];
  nlps_engine *eng = nlps_b200_create(&in.mesh, &in.solver, in.n_bounds, in.bounds, in.n_neumann, in.neumann, in.gravity,
                                      MPM_Mesh.NumberMaterials, in.mats, &in.st, 0, msg, sizeof(msg));
  if (eng == NULL) {
    fprintf(stderr, "" RED "Error in nlps_b200_create(): %s" RESET " \n", msg);
    b200_release(&in);
    return EXIT_FAILURE;
  }
  DeltaTimeStep = nlps_b200_dt(eng); /* __compute_deltat (U-Newmark-beta.c:180,438-483) */

  nlps_newmark prm;
  memset(&prm, 0, sizeof(prm));
  prm.beta = Parameters_Solver.beta_Newmark_beta;   /* U-Newmark-beta.c:146-147 */
  prm.gamma = Parameters_Solver.gamma_Newmark_beta;
  prm.tol = Parameters_Solver.TOL_Newmark_beta;     /* :171-172: rtol, atol = 100 tol */
  prm.max_iter = Parameters_Solver.MaxIter;
  prm.use_explicit_trial = quasi_static ? 0 : Parameters_Solver.Use_explicit_trial;
  prm.quasi_static = quasi_static; /* U_Static: the same loop without inertia (U-Static.c:83-322) */
  if (nlps_b200_newmark_setup(eng, &prm) != EXIT_SUCCESS) {
    fprintf(stderr, "" RED "Error in nlps_b200_newmark_setup()" RESET " \n");
    STATUS = EXIT_FAILURE;
  }

  const int Np = MPM_Mesh.NumGP, cap = nlps_b200_list_capacity(eng);
  int *counts = (int *)malloc(sizeof(int) * Np);
  int *lists = (int *)malloc(sizeof(int) * (size_t)Np * cap);
  for (int TimeStep = Parameters_Solver.InitialTimeStep; TimeStep < NumTimeStep && STATUS == EXIT_SUCCESS; TimeStep++) {
    print_step(TimeStep, NumTimeStep, DeltaTimeStep);
    if (nlps_b200_newmark_step(eng, TimeStep) != EXIT_SUCCESS) {
      fprintf(stderr, "" RED "Error in nlps_b200_newmark_step() at step %i" RESET " \n", TimeStep);
      STATUS = EXIT_FAILURE;
      break;
    }
    if (Flag_Print_Convergence) { /* __monitor (U-Newmark-beta.c:358-372) */
      nlps_newmark_stats s;
      nlps_b200_newmark_stats(eng, &s);
      print_convergence_stats(TimeStep, NumTimeStep, s.newton_iters, prm.max_iter, s.residual0, s.residual,
                              s.residual0 > 0 ? s.residual / s.residual0 : 0.0);
    }
    if (ResultsTimeStep > 0 && TimeStep % ResultsTimeStep == 0) { /* :409-411 */
      if (nlps_b200_download(eng, &in.st) != EXIT_SUCCESS || nlps_b200_get_lists(eng, counts, lists, cap) != EXIT_SUCCESS) {
        STATUS = EXIT_FAILURE;
        break;
      }
      lists_to_chains(MPM_Mesh, counts, lists, cap);
      b200_write_particle_results(MPM_Mesh, TimeStep, ResultsTimeStep);
    }
  }
  if (STATUS == EXIT_SUCCESS) STATUS = b200_finish(eng, &in, FEM_Mesh, MPM_Mesh);
  nlps_b200_destroy(eng);
  free(counts); free(lists);
  b200_release(&in);
  return STATUS;
}

int U_Newmark_Beta(Mesh FEM_Mesh, Particle MPM_Mesh, Time_Int_Params Parameters_Solver) {
  return b200_implicit_scheme(FEM_Mesh, MPM_Mesh, Parameters_Solver, 0);
}

/* PetscErrorCode U_Static(Mesh, Particle, Time_Int_Params) (Formulations/Displacements/U-Static.c:83, dispatched from
 * driver-nl-partsol.c:373-375 for `NLPS-Solver (Type=Static)`): residual f_int - f_trac - M b, tangent K, particles
 * updated in position and history only.  Linked INSTEAD OF U-Static.c. */
int U_Static(Mesh FEM_Mesh, Particle MPM_Mesh, Time_Int_Params Parameters_Solver) {
  return b200_implicit_scheme(FEM_Mesh, MPM_Mesh, Parameters_Solver, 1);
}
