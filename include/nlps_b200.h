/*
 * nlps_b200.h -- C ABI of the B200-native NL-PartSol explicit (NPC-FS) hot path.
 *
 * Plain pointers and sizes only.  Every entry point names the reference
 * interface it replaces (paths relative to nl-partsol/src of migmolper/NL-PartSol).
 * The library is libnlps_b200.so (nl-partsol_b200/csrc).  Host code of the reference
 * stays C: the scheme shim nl-partsol_b200/host/U-Verlet-b200.c flattens the
 * reference's `Mesh` / `Particle` / `Time_Int_Params` structs into these PODs and is
 * linked in place of Formulations/Displacements/U-Verlet.c.
 *
 * Conventions
 *   - all reals are IEEE fp64, all indices 32-bit int (as in the reference).
 *   - host particle arrays are the reference's own `Matrix.nV` buffers:
 *     row-major Np x cols (cols = d for vectors, T = 5 (2D) / 9 (3D) for tensors,
 *     d*d for C_ep), or bare double[Np]; see Types.h:184-283, U-Analisys.c:5-170.
 *   - adjacency is the reference's linked lists flattened to CSR in CHAIN
 *     (traversal) order: NodalLocality_0 -> ring1, NodalLocality -> ring2
 *     (Types.h:631-760, Read_GramsBox.c:334-456).
 *   - return value 0 = EXIT_SUCCESS, 1 = EXIT_FAILURE (the reference's error
 *     convention, U-Verlet.c:101-135); details via nlps_b200_last_error().
 *   - there is NO CPU fallback: every call fails loudly without a CUDA device.
 */
#ifndef NLPS_B200_H
#define NLPS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define NLPS_MAT_NEO_HOOKEAN_WRIGGERS 0 /* Constitutive/Hyperelastic/Neo-Hookean.c:38 */
#define NLPS_MAT_DRUCKER_PRAGER 1       /* Constitutive/Plasticity/Drucker-Prager.c:319 */
#define NLPS_MAT_MATSUOKA_NAKAI 2       /* Constitutive/Plasticity/Matsuoka-Nakai.c:300 */
#define NLPS_MAT_VON_MISES 3            /* Constitutive/Plasticity/Von-Mises.c:228 */
#define NLPS_MAT_HENCKY 4               /* Constitutive/Hyperelastic/Hencky.c:30 */
#define NLPS_MAT_LADE_DUNCAN 5          /* Constitutive/Plasticity/Lade-Duncan.c:290 (explicit scheme only) */

/* error codes latched on the device (first offender wins) */
#define NLPS_ERR_NONE 0
#define NLPS_ERR_NEGATIVE_JACOBIAN 2   /* U-Verlet.c:608-613 */
#define NLPS_ERR_FEW_NEIGHBOURS 3      /* LME.c:1087-1092 */
#define NLPS_ERR_SINGULAR_HESSIAN 4    /* LME.c:308-313 */
#define NLPS_ERR_NEWTON_LME 5          /* LME.c:343-350 */
#define NLPS_ERR_RETURN_MAP_DP 6       /* Drucker-Prager.c:469-482 and __eps/__kappa */
#define NLPS_ERR_RETURN_MAP_MN 7       /* Matsuoka-Nakai.c:__solver / tangent */
#define NLPS_ERR_SINGULAR_DF 8         /* TensorLib.c:829-905 (compute_adjunt) */
#define NLPS_ERR_SLAB_EXCURSION 9     /* a particle left the halo band of its slab between two migrations */
#define NLPS_ERR_SLAB_CAPACITY 10     /* migration would exceed the particle capacity of a slab */
#define NLPS_ERR_CSR_PATTERN 11       /* implicit: a particle couples two nodes outside the tangent pattern */
#define NLPS_ERR_HALO_TIMEOUT 12      /* peer-memory halo: the neighbour slab did not deliver within NLPS_HALO_TIMEOUT_S (30 s) */
#define NLPS_ERR_RETURN_MAP_VM 13     /* Von-Mises.c:__kappa / __d_kappa (negative equivalent plastic strain) */
#define NLPS_ERR_CUDA 100

/* Background mesh: the parts of `Mesh` (Types.h:631-760) the stepped path reads. */
typedef struct nlps_mesh {
  int ndim;               /* NumberDimensions (Macros.h:33-37) */
  int n_nodes;            /* Mesh.NumNodesMesh */
  const double *coords;   /* Mesh.Coordinates.nV, n_nodes x ndim */
  const int *ring1_ptr;   /* NodalLocality_0 (1 ring), CSR chain order, n_nodes+1 */
  const int *ring1_idx;
  const int *ring2_ptr;   /* NodalLocality (2 rings), CSR chain order */
  const int *ring2_idx;
  const double *h_avg;    /* Mesh.h_avg */
  double delta_x;         /* Mesh.DeltaX */
} nlps_mesh;

/* One `Load` (Types.h:296-332): Dirichlet set (ids = nodes), Neumann set
 * (ids = particles).  dir/val are dim x num_steps, value k at step s is
 * Value[k].Fx[s], active iff Dir[k*NumTimeStep+s] == 1. */
typedef struct nlps_load {
  int n_ids;
  int dim;
  const int *ids;
  const int *dir;
  const double *val;
} nlps_load;

/* The slice of `Material` (Types.h:359-458) the three in-scope laws read. */
typedef struct nlps_material {
  int type;
  double rho, E, nu;
  double reference_pressure;          /* ReferencePressure */
  double kappa_0, hardening_modulus;  /* kappa_0, Hardening_modulus */
  double plastic_strain_0;            /* Plastic_Strain_0 */
  double phi_frictional, psi_frictional; /* degrees */
  double exponent_hardening_ortiz;    /* m */
  double cohesion;
  double alpha_hardening_borja;
  double a_hardening_borja[3];
  /* Von-Mises (InOutFun/Material/Plasticity/Von-Mises.c:69-73 defaults: theta = 1, the others 0); kappa_0 is its
   * Yield-stress, hardening_modulus its Hardening-Modulus */
  double theta_hardening_voce, k_0_hardening_voce, k_inf_hardening_voce, delta_hardening_voce;
} nlps_material;

/* `Time_Int_Params` (Types.h:804-865) + the process globals the scheme reads
 * (Globals.h:16-109): gamma_LME, TOL_zero_LME, TOL_wrapper_LME, max_iter_LME,
 * TOL_Radial_Returning, Max_Iterations_Radial_Returning, Thickness_Plain_Stress. */
typedef struct nlps_solver {
  double cfl, cel;
  int initial_step, num_steps;
  double gamma_lme, tol_zero_lme, tol_wrapper_lme;
  int max_iter_lme;
  double tol_radial_returning;
  int max_iter_radial_returning;
  double thickness;
  int quirk_transposed_eigvec; /* the reference's plastic branches index the eigenvector matrix by row
                                * (Drucker-Prager.c:957-958).  -1 = default: 1 in 2D (bit-compatible with the reference),
                                * 0 in 3D (the intended column form; the compiled 3D laws do have the row form and 1
                                * reproduces them, but its result depends on the arbitrary eigenvectors of degenerate
                                * trial states -- DESIGN.md section 6, deviation 3) */
  int compute_c_ep;            /* write Phi.C_ep (needed by the implicit tangent only) */
  int shape_function;          /* the global ShapeFunctionGP (GramsShapeFun, Read_GramsShapeFun.c:84-176): NLPS_SHAPE_LME
                                * (Nodes/LME.c) or NLPS_SHAPE_ALME (Nodes/aLME.c: a d x d thermalisation metric and a
                                * cut-off ellipsoid per particle, both convected with DF^-1 at every search; 2D only, as
                                * in the reference, whose aLME.c:693-809 exits in 3D; single engine, no slabs) */
} nlps_solver;
#define NLPS_SHAPE_LME 0
#define NLPS_SHAPE_ALME 1

/* Host views of `Particle` / `Fields` (Types.h:548-623,184-283).  Any pointer may
 * be NULL in download()/upload() calls (then that field is skipped); create()
 * needs x_GC, mass, Vol_0, rho, F_n, I0, MatIdx and treats NULL as "zero /
 * identity as allocate_U_vars__Fields__ leaves it". */
typedef struct nlps_particles {
  int n; /* Particle.NumGP */
  double *x_GC, *dis, *D_dis, *vel, *acc;                 /* n x d */
  double *F_n, *F_n1, *DF, *b_e_n, *b_e_n1, *Stress;     /* n x T */
  double *C_ep;                                           /* n x d*d */
  double *J_n, *J_n1, *mass, *rho, *Vol_0, *W;            /* n */
  double *EPS_n, *EPS_n1, *Kappa_n, *Kappa_n1;            /* n */
  double *lambda; /* n x d  (Particle.lambda) */
  double *Beta;   /* n      (Particle.Beta); n x d*d with NLPS_SHAPE_ALME (Generate-One-Phase-Analysis.c:192-202) */
  int *I0, *NumberNodes, *MatIdx;                         /* n */
  double *Area_0; /* n: Particle.Phi.Area_0, the area a 3D Neumann load acts on (U-Verlet.c:847-849,
                   * U-Newmark-beta.c:1442, U-Static.c:930); 2D uses Vol_0 / Thickness_Plain_Stress and ignores it.
                   * nlps_b200_create fails for a 3D deck with Neumann loads and Area_0 == NULL. */
  double *Back_stress; /* n x 3: Particle.Phi.Back_stress (Types.h:266), the kinematic-hardening back stress of Von-Mises in
                        * PRINCIPAL components (Von-Mises.c:256-258,729-731), updated in place; NULL = zero and not
                        * written back. */
  double *Cut_off_Ellipsoid; /* n x d*d: Particle.Cut_off_Ellipsoid, the metric of the neighbour test of aLME
                              * (aLME.c:811-872); read and written with NLPS_SHAPE_ALME only, else ignored */
} nlps_particles;

typedef struct nlps_engine nlps_engine;

/* stages of one explicit step, in order */
enum nlps_stage {
  NLPS_STAGE_SEARCH = 0,        /* local_search__MeshTools__  Shape-Functions.c:31 -> LME.c:895 */
  NLPS_STAGE_P2G_MASS_DISP = 1, /* __mass_NODES, __predictor_PARTICLES, __d_displacement_NODES U-Verlet.c:166-367 */
  NLPS_STAGE_GRID_DISP = 2,     /* /M and impose_Dirichlet_Boundary_Conditions U-Verlet.c:359-363,458-526 (fused into stage 1 on the device; no-op) */
  NLPS_STAGE_KIN_STRESS = 3,    /* __update_Local_State U-Verlet.c:530-676 + Constitutive.c:18 */
  NLPS_STAGE_FORCE = 4,         /* __nodal_internal_forces U-Newmark-beta.c:1257-1374 + tractions U-Verlet.c:805-902 */
  NLPS_STAGE_GRID_ACC = 5,      /* solve_Nodal_Equilibrium nodal part U-Verlet.c:947-958 (fused into stage 4; no-op) */
  NLPS_STAGE_G2P = 6            /* G2P + compute_Explicit_Newmark_Corrector U-Verlet.c:963-1084 */
};

/* Replaces the setup the scheme functions do implicitly by receiving the
 * reference structs by value (U-Verlet.c:64-87): copies mesh, loads, materials
 * and the particle state to the device (AoS -> SoA, chains -> CSR).  gravity is
 * ndim x num_steps (gravity_field.Value[k].Fx, U-Newmark-beta.c:1539-1543) or NULL.
 * Returns NULL on failure and writes a message into err. */
nlps_engine *nlps_b200_create(const nlps_mesh *mesh, const nlps_solver *solver,
                              int n_bounds, const nlps_load *bounds,
                              int n_neumann, const nlps_load *neumann,
                              const double *gravity, int n_materials,
                              const nlps_material *materials,
                              const nlps_particles *state, int device,
                              char *err, int err_len);
void nlps_b200_destroy(nlps_engine *e);

/* initialize__LME__ (LME.c:45-173) phases 2-3, given I0: activate nodes,
 * first neighbour lists (Beta = 0 => infinite radius), Beta, Newton for lambda. */
int nlps_b200_initialize_lme(nlps_engine *e);

/* One iteration of the U_Verlet time loop (U-Verlet.c:89-159, intended form). */
int nlps_b200_step(nlps_engine *e, int time_step);
/* `count` consecutive steps; the error flag is polled once at the end. */
int nlps_b200_run(nlps_engine *e, int first_step, int count);
/* nlps_b200_run bracketed by CUDA events on the engine's stream; *ms = device time of the steps. */
int nlps_b200_timed_run(nlps_engine *e, int first_step, int count, double *ms);
/* One stage (stage-by-stage parity tests). */
int nlps_b200_stage(nlps_engine *e, int stage, int time_step);

/* D2H of particle fields into the caller's (reference-owned) buffers before
 * particle_results_vtk__InOutFun__ (U-Verlet.c:1088-1227); never reallocates. */
int nlps_b200_download(nlps_engine *e, nlps_particles *out);
int nlps_b200_upload(nlps_engine *e, const nlps_particles *in);

/* Output overlapped with stepping (the reference stops the time loop for every
 * particle_results_vtk__InOutFun__, U-Verlet.c:1088-1227 / InOutFun/Outputs/WriteVtk.c:95-268):
 *   nlps_b200_run_async      enqueue `count` steps and return at once;
 *   nlps_b200_sync           wait for everything enqueued, poll the error flag (what nlps_b200_run ends with);
 *   nlps_b200_download_begin snapshot the fields of the current step on the device (stream-ordered: steps enqueued
 *                            afterwards do not disturb it);
 *   nlps_b200_download_end   copy the snapshot into `out` of the matching begin and wait for the copy only.
 * Typical loop: run_async(to the output step) ... sync; download_begin; run_async(next chunk); download_end;
 * write the files while the GPU steps on; sync.  `out` must stay valid between begin and end.  When the snapshot
 * buffer cannot be had, begin downloads synchronously and end is a no-op. */
int nlps_b200_run_async(nlps_engine *e, int first_step, int count);
int nlps_b200_sync(nlps_engine *e);
int nlps_b200_download_begin(nlps_engine *e, nlps_particles *out);
int nlps_b200_download_end(nlps_engine *e);

/* Nodal arrays of the last step in full-grid indexing, n_nodes x ndim, zero on
 * inactive nodes.  which: 0 lumped mass (per DOF, d identical copies as
 * U-Verlet.c:216), 1 D_Displacement, 2 Forces, 3 Acceleration, 4 Reactions. */
int nlps_b200_get_nodal(nlps_engine *e, int which, double *out);
int nlps_b200_get_active(nlps_engine *e, unsigned char *out); /* Mesh.ActiveNode */
/* Particle.ListNodes expanded in chain order: lists is n x cap (-1 padded). */
int nlps_b200_list_capacity(nlps_engine *e);
int nlps_b200_get_lists(nlps_engine *e, int *counts, int *lists, int cap);

int nlps_b200_last_error(nlps_engine *e, int *code, int *particle);
double nlps_b200_dt(nlps_engine *e); /* U_DeltaT__SolversLib__ Courant.c:6 */

/* Per-kernel device times (CUDA events on the engine's stream), accumulated
 * since the last reset.  names/ms/launches have room for `cap` entries;
 * returns the number of kernels. */
int nlps_b200_profile(nlps_engine *e, int enable);
int nlps_b200_kernel_times(nlps_engine *e, int cap, const char **names, double *ms, int *launches);
void nlps_b200_reset_kernel_times(nlps_engine *e);
long long nlps_b200_launch_count(nlps_engine *e);

/* The whole scheme call with HOST buffers (what U_Verlet does for the driver):
 * create + initialize (optional) + steps [initial_step, num_steps) + download
 * every `results_every` steps (callback may be NULL) + destroy. */
typedef void (*nlps_results_cb)(int time_step, void *user);
int nlps_b200_u_verlet(const nlps_mesh *mesh, const nlps_solver *solver,
                       int n_bounds, const nlps_load *bounds, int n_neumann,
                       const nlps_load *neumann, const double *gravity,
                       int n_materials, const nlps_material *materials,
                       nlps_particles *state, int run_initialize,
                       int results_every, nlps_results_cb cb, void *user,
                       int device);

/* Scalable setup: O(N) replacement of get_sourrounding_elements /
 * fill_nodal_locality / compute_nodal_distance_local / mesh_size
 * (Read_GramsBox.c:293-565) producing the same chain orders.  Host code.
 * Two-call protocol: sizes first (idx pointers NULL), then fill. */
int nlps_b200_build_locality(int ndim, int n_nodes, int n_elems, int nodes_per_elem,
                             const int *connectivity, const double *coords,
                             int *ring1_ptr, int *ring1_idx, int *ring2_ptr,
                             int *ring2_idx, double *h_avg, double *delta_x);

/* Stress_integration__Constitutive__ (Constitutive/Constitutive.c:18-258) over an array of
 * independent material points (host AoS rows of T doubles); status[p] = 0 or an NLPS_ERR_* code. */
int nlps_b200_stress_points(int ndim, const nlps_material *material, double tol_radial,
                            int max_iter_radial, int quirk_transposed_eigvec, int n,
                            const double *DF, const double *F_n1, const double *J_n1,
                            const double *b_e_n, const double *eps_n, const double *kappa_n,
                            double *stress, double *b_e_n1, double *eps_n1, double *kappa_n1,
                            double *W, double *C_ep, int *status, int device);

/* ---------------------------------------------------------------------------
 * Multi-GPU: spatial slabs.  The reference is single-process (OpenMP only,
 * driver-nl-partsol.c:192); this is new surface, one engine per GPU / process.
 *
 * Every slab engine holds the whole (static) background mesh and the particles whose closest
 * node I0 lies between its two cuts along `axis`.  Per step three halo exchanges with the two
 * neighbour slabs over the nodes within `band_cells` cells of a cut: cell occupancy (feeds
 * Mesh.ActiveNode, LME.c:949-965), lumped mass + momentum sums (U-Verlet.c:166-225,301-367) and
 * force sums (U-Newmark-beta.c:1257-1374); both sides then divide / apply the Dirichlet set
 * identically, so no return broadcast is needed.  Every `migrate_every` steps particles whose I0
 * crossed a cut move to the neighbour slab (full state row).
 *
 * Transport: NCCL (ncclSend/ncclRecv grouped on the engine's stream) or a caller-supplied
 * exchange function (tests, other fabrics).  Host code distributes the NCCL id itself
 * (torch.distributed, MPI, a file ...). */
typedef struct nlps_msg {
  int peer;               /* slab rank of the other side */
  const void *send;       /* device pointers */
  unsigned long long send_bytes;
  void *recv;
  unsigned long long recv_bytes;
} nlps_msg;
/* post all messages on CUDA stream `cuda_stream` (a cudaStream_t); 0 on success */
typedef int (*nlps_exchange_fn)(void *user, int n_msgs, const nlps_msg *msgs, void *cuda_stream);
typedef struct nlps_comm nlps_comm;
int nlps_b200_comm_unique_id(char id[128]); /* ncclGetUniqueId */
nlps_comm *nlps_b200_comm_create_nccl(const char id[128], int rank, int world, int device);
nlps_comm *nlps_b200_comm_create_custom(int rank, int world, nlps_exchange_fn fn, void *user);
void nlps_b200_comm_destroy(nlps_comm *c);
/* helpers for custom transports */
int nlps_b200_memcpy_d2d(void *dst, const void *src, unsigned long long bytes, void *cuda_stream);
int nlps_b200_stream_sync(void *cuda_stream);

typedef struct nlps_slab {
  int rank, world;
  int axis;               /* slab axis (0..ndim-1) */
  const double *cuts;     /* world-1 ascending cut coordinates (between node layers) */
  int band_cells;         /* halo half-width in cells of size delta_x (>= 4; 0 = default 6) */
  int migrate_every;      /* steps between migrations (0 = default 10) */
  double capacity_factor; /* particle capacity / initial count (0 = default 1.3) */
  int n_global;           /* global particle count (ids are 0..n_global-1) */
  const int *global_id;   /* id of every row of `state`, or NULL: row index == global id */
  int node_id_offset;     /* sub-mesh slabs: global node id = local node id + node_id_offset (the sub-mesh is
                             a contiguous id range of the global mesh, e.g. whole rows / planes of a structured
                             grid); 0 when every slab holds the whole mesh.  Used when particles migrate. */
  nlps_comm *comm;
} nlps_slab;

/* Host planning: slab axis (longest extent of the cloud when axis < 0) and world-1 cuts that
 * balance the particle counts, each placed midway between two node layers. */
int nlps_b200_slab_cuts(const nlps_mesh *mesh, int n, const int *I0, int world, int axis,
                        int *axis_out, double *cuts_out);
/* Which slab owns a particle whose closest node is I0 (host). */
int nlps_b200_slab_owner(const nlps_mesh *mesh, int axis, int world, const double *cuts, int I0);
/* Node ids within band_cells*delta_x of cut (ascending ids): the halo both neighbours exchange.
 * ids may be NULL (count only). */
int nlps_b200_slab_halo_nodes(const nlps_mesh *mesh, int axis, double cut, int band_cells, int *ids);

/* nlps_b200_create for one slab: keeps the rows of `state` this slab owns. */
nlps_engine *nlps_b200_create_slab(const nlps_mesh *mesh, const nlps_solver *solver,
                                   int n_bounds, const nlps_load *bounds,
                                   int n_neumann, const nlps_load *neumann,
                                   const double *gravity, int n_materials,
                                   const nlps_material *materials,
                                   const nlps_particles *state, const nlps_slab *slab,
                                   int device, char *err, int err_len);
int nlps_b200_local_count(nlps_engine *e);
/* which data plane carries the per-step halo sums of this slab engine (a static string) */
const char *nlps_b200_transport(nlps_engine *e);
/* compact download: rows 0..local_count-1 of `out` (out->n >= local_count) and their global ids */
int nlps_b200_download_local(nlps_engine *e, nlps_particles *out, int *ids);
/* move particles that crossed a cut now (also runs every migrate_every steps inside run) */
int nlps_b200_migrate(nlps_engine *e);
long long nlps_b200_migrated_count(nlps_engine *e); /* particles received so far */
/* The scheme call for one slab; `state` rows are indexed by global id on return (rows of other
 * slabs untouched) when slab->global_id == NULL; otherwise compact (download_local order): state->n
 * becomes the number of particles the slab holds at the end and ids_out (may be NULL, room for the
 * input state->n... capacity rows) receives their global ids. */
int nlps_b200_u_verlet_slab(const nlps_mesh *mesh, const nlps_solver *solver,
                            int n_bounds, const nlps_load *bounds, int n_neumann,
                            const nlps_load *neumann, const double *gravity,
                            int n_materials, const nlps_material *materials,
                            nlps_particles *state, const nlps_slab *slab, int *ids_out,
                            int run_initialize, int results_every, nlps_results_cb cb, void *user,
                            int device);

/* ---------------------------------------------------------------------------
 * Implicit scheme: "Newmark-beta-Finite-Strains" -> PetscErrorCode U_Newmark_Beta(Mesh, Particle,
 * Time_Int_Params) (Formulations/Displacements/U-Newmark-beta.c:130-425).  The PETSc objects of
 * the reference (Vec/Mat/IS/SNES/KSP/PC, :220-425) are replaced by device vectors over the active
 * nodes, a block-CSR tangent and a hand-written Jacobi-PCG; the nonlinear driver is Newton with
 * step halving.  Tangents: Neo-Hookean-Wriggers (Neo-Hookean.c:89-141; symmetric, Jacobi-PCG) and the spectral
 * elastoplastic tangent of Drucker-Prager / Matsuoka-Nakai (Elastoplastic-Tangent-Matrix.c:42-160; unsymmetric,
 * Jacobi-BiCGStab).  Slab engines (nlps_b200_create_slab) run the Neo-Hookean operator over several GPUs: every slab
 * assembles the tangent of its own particles, the Krylov vectors are summed over the halo-band nodes each iteration
 * and the dot products over the slabs; the elastoplastic tangents are single-slab. */
typedef struct nlps_newmark {
  double beta, gamma;      /* Time_Int_Params.beta_Newmark_beta / gamma_Newmark_beta */
  double tol;              /* TOL_Newmark_beta: rtol; atol = 100 * tol (U-Newmark-beta.c:171-172) */
  int max_iter;            /* MaxIter (Newton) */
  int use_explicit_trial;  /* Use_explicit_trial (:879-957) */
  double pcg_rtol;         /* |r| <= pcg_rtol |b|   (0 = 1e-8; PETSc's default would be 1e-5) */
  int pcg_max_iter;        /* 0 = 10000 (PETSc default) */
  int quasi_static;        /* 1: U_Static (Formulations/Displacements/U-Static.c:83-322) -- the same loop without inertia:
                            * residual f_int - f_trac - M b, tangent K only, particles updated in position and history,
                            * velocities and accelerations untouched; beta / gamma are ignored */
} nlps_newmark;
typedef struct nlps_newmark_stats {
  int newton_iters;        /* of the last step */
  int n_rows, nnz_blocks;  /* block rows (active nodes) and d x d blocks of the last pattern */
  long long pcg_iters_total, assemblies_total, residual_evals_total;
  double residual0, residual; /* |R| before / after the last step's Newton loop */
  double ms_assemble, ms_pcg, ms_residual; /* accumulated device times */
} nlps_newmark_stats;
int nlps_b200_newmark_setup(nlps_engine *e, const nlps_newmark *prm);
int nlps_b200_newmark_step(nlps_engine *e, int time_step);  /* one pass of the loop :192-425 */
int nlps_b200_newmark_run(nlps_engine *e, int first_step, int count);
int nlps_b200_newmark_stats(nlps_engine *e, nlps_newmark_stats *out);
/* stage-level entry points (parity tests): search + lumped mass + v_n, a_n + pattern + initial guess */
int nlps_b200_newmark_begin(nlps_engine *e, int time_step);
/* which: 0 v_n, 1 a_n, 2 dU (current iterate), 3 residual; out is n_nodes x ndim (full grid) */
int nlps_b200_newmark_get(nlps_engine *e, int which, double *out);
/* __lagrangian_evaluation (:970-1050) at dU (n_nodes x ndim; NULL = current iterate) */
int nlps_b200_newmark_residual(nlps_engine *e, int time_step, const double *dU, double *R);
/* __jacobian_evaluation (:1646-1830) at the state of the last residual evaluation, as block CSR over
 * the active nodes, without alpha_1 M and without the Dirichlet rows/columns (applied in the
 * operator).  Two calls: sizes first (arrays NULL), then fill. */
int nlps_b200_newmark_tangent(nlps_engine *e, int *n_rows, int *nnz_blocks, int *row_nodes,
                              int *row_ptr, int *col_nodes, double *vals);
int nlps_b200_u_newmark_beta(const nlps_mesh *mesh, const nlps_solver *solver,
                             const nlps_newmark *newmark, int n_bounds, const nlps_load *bounds,
                             int n_neumann, const nlps_load *neumann, const double *gravity,
                             int n_materials, const nlps_material *materials,
                             nlps_particles *state, int run_initialize, int results_every,
                             nlps_results_cb cb, void *user, int device);

/* Engines take their device memory from the device's stream-ordered pool and leave it cached there on
 * destroy (a second scheme call in the same process then skips cudaMalloc/cudaFree); this returns the cached
 * memory to the driver.  NLPS_POOL=0 in the environment disables the pool. */
int nlps_b200_trim(int device);

const char *nlps_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NLPS_B200_H */
