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
 *      particle_results_vtk__InOutFun__ / nodal_results_vtk__InOutFun__ call (every ResultsTimeStep steps,
 *      U-Verlet.c:1088-1227); the copy and the writers of output step k overlap the steps after k.
 *
 * Error convention as the reference: EXIT_SUCCESS / EXIT_FAILURE, RED message on stderr.
 */
#include "b200_flatten.h"
#include "b200_vtk_binary.h" /* NLPS_B200_VTK_BINARY=1: binary twin of the reference's VTK writer */

#include <pthread.h>

double DeltaTimeStep; /* defined by U-Verlet.c:3 in the reference; other TUs reference it */

/*
 * Several GPUs from the C host: NLPS_B200_GPUS=n (beside --OPENMP-CORES, driver-nl-partsol.c:148-153).
 * One host thread per device; the cloud is cut into n spatial slabs along its longest extent (nlps_b200_slab_cuts,
 * particle-count quantiles); every thread runs the scheme call of ITS slab (nlps_b200_u_verlet_slab: create on device r,
 * halo sums with the two neighbour slabs every step, particle migration every 10 steps over NCCL) and writes the rows
 * of the particles it holds into the reference's own field buffers, which all threads share (rows are indexed by the
 * caller's particle id, the slabs hold disjoint sets).  At a results step the threads meet at a barrier and thread 0
 * calls the reference's writer.  No torch, no MPI: the NCCL id is made by thread 0 and read by the others.
 */
typedef struct slab_job {
  int rank, world, axis, results_every, status;
  const b200_inputs *in;
  const double *cuts;
  const char *nccl_id;
  Particle MPM_Mesh;
  pthread_barrier_t *bar;
} slab_job;

static void slab_results_cb(int time_step, void *user) {
  slab_job *j = (slab_job *)user;
  pthread_barrier_wait(j->bar); /* every slab has written the rows it holds */
  if (j->rank == 0) b200_write_particle_results(j->MPM_Mesh, time_step, j->results_every);
  pthread_barrier_wait(j->bar); /* the buffers are free again */
}

static void *slab_thread(void *arg) {
  slab_job *j = (slab_job *)arg;
  nlps_comm *comm = nlps_b200_comm_create_nccl(j->nccl_id, j->rank, j->world, j->rank);
  if (comm == NULL) {
    fprintf(stderr, "" RED "Error in nlps_b200_comm_create_nccl() (slab %i of %i)" RESET " \n", j->rank, j->world);
    j->status = EXIT_FAILURE;
    return NULL;
  }
  nlps_slab sl;
  memset(&sl, 0, sizeof(sl));
  sl.rank = j->rank; sl.world = j->world; sl.axis = j->axis; sl.cuts = j->cuts;
  sl.n_global = j->in->st.n; sl.global_id = NULL; sl.node_id_offset = 0; sl.comm = comm;
  nlps_particles st = j->in->st; /* the shared buffers of the reference: this slab touches its own rows only */
  j->status = nlps_b200_u_verlet_slab(&j->in->mesh, &j->in->solver, j->in->n_bounds, j->in->bounds, j->in->n_neumann,
                                      j->in->neumann, j->in->gravity, j->MPM_Mesh.NumberMaterials, j->in->mats, &st, &sl, NULL, 0,
                                      j->results_every, slab_results_cb, j, j->rank);
  nlps_b200_comm_destroy(comm);
  return NULL;
}

static int u_verlet_slabs(int ngpu, b200_inputs *in, Particle MPM_Mesh, Time_Int_Params Parameters_Solver) {
  int axis = 0, STATUS = EXIT_SUCCESS;
  double *cuts = (double *)calloc(ngpu, sizeof(double));
  char id[128];
  if (nlps_b200_slab_cuts(&in->mesh, in->st.n, in->st.I0, ngpu, -1, &axis, cuts) != EXIT_SUCCESS ||
      nlps_b200_comm_unique_id(id) != EXIT_SUCCESS) {
    fprintf(stderr, "" RED "Error in U_Verlet() [B200]: cannot plan %i slabs (cuts / NCCL id)" RESET " \n", ngpu);
    free(cuts);
    return EXIT_FAILURE;
  }
  DeltaTimeStep = Parameters_Solver.CFL * in->mesh.delta_x / Parameters_Solver.Cel; /* Courant.c:6-55 */
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, NULL, ngpu);
  slab_job *jobs = (slab_job *)calloc(ngpu, sizeof(slab_job));
  pthread_t *th = (pthread_t *)calloc(ngpu, sizeof(pthread_t));
  for (int r = 0; r < ngpu; r++) {
    jobs[r].rank = r; jobs[r].world = ngpu; jobs[r].axis = axis; jobs[r].results_every = ResultsTimeStep;
    jobs[r].in = in; jobs[r].cuts = cuts; jobs[r].nccl_id = id; jobs[r].MPM_Mesh = MPM_Mesh; jobs[r].bar = &bar;
    jobs[r].status = EXIT_FAILURE;
    pthread_create(&th[r], NULL, slab_thread, &jobs[r]);
  }
  for (int r = 0; r < ngpu; r++) {
    pthread_join(th[r], NULL);
    if (jobs[r].status != EXIT_SUCCESS) STATUS = EXIT_FAILURE;
  }
  pthread_barrier_destroy(&bar);
  free(jobs); free(th); free(cuts);
  return STATUS;
}

int U_Verlet(Mesh FEM_Mesh, Particle MPM_Mesh, Time_Int_Params Parameters_Solver) {
  const int NumTimeStep = Parameters_Solver.NumTimeStep;
  const int Np = MPM_Mesh.NumGP;
  int STATUS = EXIT_SUCCESS;

  if (strcmp(ShapeFunctionGP, "LME") != 0 && (strcmp(ShapeFunctionGP, "aLME") != 0 || NumberDimensions != 2)) {
    fprintf(stderr, "" RED "Error in U_Verlet() [B200]: only GramsShapeFun (Type=LME) and, in 2D, (Type=aLME) are supported" RESET " \n");
    return EXIT_FAILURE;
  }
  b200_inputs in;
  if (b200_flatten(&in, FEM_Mesh, MPM_Mesh, Parameters_Solver) != EXIT_SUCCESS) return EXIT_FAILURE;

  const char *ng = getenv("NLPS_B200_GPUS");
  if (ng != NULL && atoi(ng) > 1) {
    /* (the neighbour chains keep the state initialise_shapefun__MeshTools__ left: the slab engines are gone when the
     * scheme call returns; NumberNodes and every field are those of the last step) */
    STATUS = u_verlet_slabs(atoi(ng), &in, MPM_Mesh, Parameters_Solver);
    b200_release(&in);
    return STATUS;
  }

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
  b200_nodal nodal;
  memset(&nodal, 0, sizeof(nodal));
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
      b200_nodal_write(&nodal, FEM_Mesh, pending_out, ResultsTimeStep);
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
      if (nlps_b200_download_begin(eng, &in.st) != EXIT_SUCCESS ||
          b200_nodal_capture(eng, &nodal, FEM_Mesh.NumNodesMesh, NumberDimensions) != EXIT_SUCCESS) {
        STATUS = EXIT_FAILURE;
        break;
      }
      pending_out = TimeStep - 1;
    }
    print_Status("DONE !!!", TimeStep - 1);
  }
  if (STATUS == EXIT_SUCCESS && pending_out >= 0) {
    if (nlps_b200_download_end(eng) != EXIT_SUCCESS) STATUS = EXIT_FAILURE;
    else {
      b200_write_particle_results(MPM_Mesh, pending_out, ResultsTimeStep);
      b200_nodal_write(&nodal, FEM_Mesh, pending_out, ResultsTimeStep);
    }
  }
  b200_nodal_release(&nodal);

  if (STATUS == EXIT_SUCCESS) STATUS = b200_finish(eng, &in, FEM_Mesh, MPM_Mesh);

  nlps_b200_destroy(eng);
  b200_release(&in);
  return STATUS;
}
