/*
 * U-Verlet-b200.c -- drop-in replacement for the reference's explicit scheme function.
 *
 * Same symbol, same by-value struct arguments and same return convention as
 *     int U_Verlet(Mesh FEM_Mesh, Particle MPM_Mesh, Time_Int_Params Parameters_Solver)
 * (Formulations/Displacements/U-Verlet.c:64, dispatched from driver-nl-partsol.c:362 when the deck
 * says `NLPS-Solver (Type=NPC-FS)`).  This translation unit is compiled AGAINST THE REFERENCE'S OWN
 * HEADERS and linked INSTEAD OF U-Verlet.c; everything else of the reference (driver, deck parser,
 * mesh / particle generation, VTK writers) stays untouched C.  It is host glue only:
 *
 *   1. flatten the reference's linked lists (NodalLocality_0, NodalLocality) to CSR in chain order,
 *      the Dirichlet / Neumann `Load` tables and the gravity `Load` to dense (dim x NumTimeStep)
 *      arrays, the `Material` structs to nlps_material PODs;
 *   2. hand the reference's own field buffers (Matrix.nV) to libnlps_b200 through the C ABI of
 *      include/nlps_b200.h -- the engine never frees or reallocates them (driver ownership,
 *      driver-nl-partsol.c:575-660);
 *   3. run the steps on the B200, copying fields back into those buffers before each
 *      particle_results_vtk__InOutFun__ call (every ResultsTimeStep steps, U-Verlet.c:1088-1227); the copy and the
 *      writer of output step k overlap the steps after k.
 *
 * Error convention as the reference: EXIT_SUCCESS / EXIT_FAILURE, RED message on stderr.
 */
#include "b200_flatten.h"
#include "b200_vtk_binary.h" /* NLPS_B200_VTK_BINARY=1: binary twin of the reference's VTK writer */

double DeltaTimeStep; /* defined by U-Verlet.c:3 in the reference; other TUs reference it */

int U_Verlet(Mesh FEM_Mesh, Particle MPM_Mesh, Time_Int_Params Parameters_Solver) {
  const int NumTimeStep = Parameters_Solver.NumTimeStep;
  const int Np = MPM_Mesh.NumGP;
  int STATUS = EXIT_SUCCESS;

  if (strcmp(ShapeFunctionGP, "LME") != 0) {
    fprintf(stderr, "" RED "Error in U_Verlet() [B200]: only GramsShapeFun (Type=LME) is supported" RESET " \n");
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
  DeltaTimeStep = nlps_b200_dt(eng);

  /* Output overlapped with stepping: at an output step the fields are snapshotted on the device
   * (nlps_b200_download_begin), the next chunk of steps is enqueued, and only then is the snapshot copied into the
   * reference's buffers and particle_results_vtk__InOutFun__ run -- the (slow, ASCII) writer of step k works while
   * the B200 computes the steps after k.  The writer reads Phi and I0 / MatIdx only (WriteVtk.c:95-268), not the
   * neighbour chains, which are rebuilt once at the end (b200_finish). */
  int pending_out = -1;
  for (int TimeStep = Parameters_Solver.InitialTimeStep; TimeStep < NumTimeStep && STATUS == EXIT_SUCCESS;) {
    /* run up to (and including) the next output step in one go */
    int next_out = TimeStep;
    if (ResultsTimeStep > 0) next_out = ((TimeStep + ResultsTimeStep - 1) / ResultsTimeStep) * ResultsTimeStep;
    int count = (ResultsTimeStep > 0) ? next_out - TimeStep + 1 : NumTimeStep - TimeStep;
    if (TimeStep + count > NumTimeStep) count = NumTimeStep - TimeStep;
    print_step(TimeStep, NumTimeStep, DeltaTimeStep);
    if (nlps_b200_run_async(eng, TimeStep, count) != EXIT_SUCCESS) {
      fprintf(stderr, "" RED "Error in nlps_b200_run_async() at steps [%i,%i)" RESET " \n", TimeStep, TimeStep + count);
      STATUS = EXIT_FAILURE;
      break;
    }
    if (pending_out >= 0) {
      if (nlps_b200_download_end(eng) != EXIT_SUCCESS) { STATUS = EXIT_FAILURE; break; }
      b200_write_particle_results(MPM_Mesh, pending_out, ResultsTimeStep);
      pending_out = -1;
    }
    if (nlps_b200_sync(eng) != EXIT_SUCCESS) {
      fprintf(stderr, "" RED "Error in nlps_b200_run() at steps [%i,%i)" RESET " \n", TimeStep, TimeStep + count);
      STATUS = EXIT_FAILURE;
      break;
    }
    TimeStep += count;
    if (ResultsTimeStep > 0 && (TimeStep - 1) % ResultsTimeStep == 0) {
      /* output_selector (U-Verlet.c:1088-1227): vtk results read the host Fields */
      if (nlps_b200_download_begin(eng, &in.st) != EXIT_SUCCESS) { STATUS = EXIT_FAILURE; break; }
      pending_out = TimeStep - 1;
    }
    print_Status("DONE !!!", TimeStep - 1);
  }
  if (STATUS == EXIT_SUCCESS && pending_out >= 0) {
    if (nlps_b200_download_end(eng) != EXIT_SUCCESS) STATUS = EXIT_FAILURE;
    else b200_write_particle_results(MPM_Mesh, pending_out, ResultsTimeStep);
  }

  if (STATUS == EXIT_SUCCESS) STATUS = b200_finish(eng, &in, FEM_Mesh, MPM_Mesh);

  nlps_b200_destroy(eng);
  b200_release(&in);
  return STATUS;
}
