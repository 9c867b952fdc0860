// nlps_implicit.inl -- implicit Newmark-beta, finite strains, included at the end of
// nlps_engine.cu.  Replaces PetscErrorCode U_Newmark_Beta(Mesh, Particle, Time_Int_Params)
// (Formulations/Displacements/U-Newmark-beta.c:130-425) below the C ABI:
//
//   K1   lumped mass, nodal v_n and a_n            :528-597, :615-696   two passes of the fused LME/P2G kernel
//   K6   initial guess + Dirichlet                 :879-957             k_imp_guess
//   K2/3 residual = internal - traction + inertia  :970-1050            the explicit kinematics/stress/force kernels
//                                                                       (rho left alone) + k_imp_residual
//   K5   tangent                                   :1568-1632, :1646-1830   block-CSR over the active nodes:
//        pattern from a static coupling adjacency filtered by ActiveNode, values by one warp per particle
//        (Neo-Hookean.c:89-141), alpha_1 M and the Dirichlet rows/columns applied inside the operator
//   K6   linear solve (PETSc KSP + PCJACOBI, :323-334) -> hand-written Jacobi-PCG on the device, all scalars
//        kept in device memory (deterministic two-stage reductions); Newton with step halving on the host
//        (SNES NEWTONLS :270-344: compare converged states, never iteration counts -- SURVEY 8c)
//   G3/K4 kinetic increments, FLIP update, roll    :1859-2072           k_g2p_implicit + pointer swap
//
// Scope: Neo-Hookean-Wriggers tangent (BASELINE configs[4]); single slab.

static const int IMP_NPART = 296;   // blocks of the vector kernels = length of their partial-sum arrays (2 x 148 SMs)
static const int IMP_NSPMV = 1184;  // blocks of the SpMV (8 x 148 SMs x 8 warps: the row gathers need the warps to hide latency)

struct ImplicitCtx {
  nlps_newmark prm{};
  double a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0;
  int *cpl_ptr = nullptr, *cpl_idx = nullptr;  // static coupling adjacency: nodes that can share a particle
  size_t cpl_total = 0;
  int *row_ptr = nullptr, *cols = nullptr, *dummy_a = nullptr, *dummy_b = nullptr, *tops = nullptr;
  double* vals = nullptr;
  size_t cap_blocks = 0;
  ulonglong2 *packed = nullptr, *scan_blk = nullptr;
  double *Vn = nullptr, *An = nullptr, *dU = nullptr, *R = nullptr, *delta = nullptr, *trial = nullptr, *Rt = nullptr;
  double *r = nullptr, *z = nullptr, *p = nullptr, *Ap = nullptr, *diag = nullptr;
  double *bv = nullptr, *bs = nullptr, *bt = nullptr, *by = nullptr, *brh = nullptr;  // BiCGStab: v, s, t, y, r^
  double* part2 = nullptr;   // BiCGStab partials: rho[2], rr (IMP_NPART each), then rv[2], ts[2], tt[2] (IMP_NSPMV each)
  int plastic = 0;           // some material has an elastoplastic tangent: unsymmetric operator
  // several slabs (SURVEY 8e "Implicit"): every slab assembles the tangent of ITS particles (K = sum of the slabs' K);
  // vectors are consistent (identical on the band nodes both slabs hold).  y = K x: local product, then the band sums of
  // the explicit scheme's force exchange; dot products count every node once (own) and are summed over the slabs.
  unsigned char* own = nullptr;  // per active rank: this slab owns the node (its coordinate lies between the slab's cuts)
  double* pair = nullptr;        // [2][2 x IMP_NPART]: r.z | r.r partials by iteration parity (one all-reduce for both)
  double* ar_tab = nullptr;      // scratch of comm_allreduce_sum for transports without a collective
  unsigned char* fx = nullptr;
  double* part = nullptr;    // [4][IMP_NPART]: 0-1 rz (ping-pong), 2 rr, 3 scratch (|R|^2); then pAp[IMP_NSPMV]
  double* h_part = nullptr;  // pinned
  std::vector<void*> allocs;
  int newton_iters = 0;
  long long pcg_iters = 0, assemblies = 0, residual_evals = 0;
  double res0 = 0, res = 0;
  double ms_assemble = 0, ms_pcg = 0, ms_residual = 0;
};

static void implicit_free(nlps_engine* e) {
  if (!e->imp) return;
  for (void* p : e->imp->allocs) pool_free(p, e->stream);
  if (e->imp->h_part) cudaFreeHost(e->imp->h_part);
  delete e->imp;
  e->imp = nullptr;
}

template <typename Tp>
static int imp_alloc(nlps_engine* e, Tp** p, size_t n) {
  void* q = nullptr;
  cudaError_t st = pool_malloc(&q, std::max<size_t>(n, 1) * sizeof(Tp), e->stream);
  if (st != cudaSuccess) {
    fprintf(stderr, "nlps_b200 (implicit): cudaMalloc(%zu) failed: %s\n", n * sizeof(Tp), cudaGetErrorString(st));
    return 1;
  }
  cudaMemsetAsync(q, 0, std::max<size_t>(n, 1) * sizeof(Tp), e->stream);
  e->imp->allocs.push_back(q);
  *p = (Tp*)q;
  return 0;
}

// ---------------------------------------------------------------------------
// nodal kernels (compact indexing by active rank t: dof = t*D + i)
__device__ __forceinline__ unsigned bc_bits(const BcDev& bc, int A, int step, int D, double* val) {
  unsigned fx = 0;
  for (int q = bc.node_ptr[A]; q < bc.node_ptr[A + 1]; q++) {
    const int b = bc.node_bnd[q];
    for (int k = 0; k < bc.bnd_dim[b] && k < D; k++) {
      const size_t o = ((size_t)b * bc.maxdim + k) * bc.nsteps + step;
      if (bc.dir[o] == 1) {
        fx |= 1u << k;
        if (val) val[k] = bc.val[o];
      }
    }
  }
  return fx;
}

// v_A or a_A = sum m N (.) / M on free dofs, 0 on restricted ones (U-Newmark-beta.c:615-696: contributions to
// restricted dofs are dropped by VEC_IGNORE_NEGATIVE_INDICES); first pass also records the restricted-dof masks
template <int D>
__global__ void __launch_bounds__(128) k_imp_nodal(GridDev G, BcDev bc, int step, int first, double* out, unsigned char* fxr) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *G.n_active) return;
  const int A = G.act_list[t];
  const unsigned fx = bc_bits(bc, A, step, D, nullptr);
  if (first) { fxr[t] = (unsigned char)fx; G.fixed[A] = (unsigned char)fx; }
  const double M = G.M[A];
#pragma unroll
  for (int i = 0; i < D; i++)  // (M = 0: a node a neighbour slab's particles activated beyond the exchanged band; nothing reads it)
    out[(size_t)t * D + i] = (((fx >> i) & 1u) || !(M > 0.0)) ? 0.0 : G.MOM[(size_t)A * D + i] / M;
}

// __form_initial_guess (U-Newmark-beta.c:879-957)
template <int D>
__global__ void __launch_bounds__(128) k_imp_guess(GridDev G, BcDev bc, int step, double dt, int explicit_trial,
                                                   const double* Vn, const double* An, double* dU) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *G.n_active) return;
  const int A = G.act_list[t];
  double val[3] = {0.0, 0.0, 0.0};
  const unsigned fx = bc_bits(bc, A, step, D, val);
#pragma unroll
  for (int i = 0; i < D; i++) {
    double u = explicit_trial ? dt * Vn[(size_t)t * D + i] + 0.5 * (dt * dt) * An[(size_t)t * D + i] : 0.0;
    if ((fx >> i) & 1u) u = val[i];
    dU[(size_t)t * D + i] = u;
  }
}

// nodal increment -> the per-node record the kinematics kernel stages
template <int D>
__global__ void __launch_bounds__(128) k_imp_set_dU(GridDev G, const double* dU) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *G.n_active) return;
  const int A = G.act_list[t];
#pragma unroll
  for (int i = 0; i < D; i++) G.UA[(size_t)A * 2 * NS<D>::X + i] = dU[(size_t)t * D + i];
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  __syncthreads();
  return s;  // valid in thread 0
}
__device__ __forceinline__ double sum_partials(const double* part, int n = IMP_NPART) {  // fixed order: identical in every block
  double s = 0.0;
  for (int i = 0; i < n; i++) s += part[i];
  return s;
}

// residual (U-Newmark-beta.c:970-1050): G.F holds (-internal + traction) from the force stage; + inertia (:1519-1557);
// restricted dofs carry none.  part_out[block] = partial sum of R^2.
template <int D>
__global__ void __launch_bounds__(256) k_imp_residual(GridDev G, const double* grav, int nsteps, int step, double a1, double a2,
                                                      double a3, const double* dU, const double* Vn, const double* An,
                                                      const unsigned char* fxr, double* R, double* part_out,
                                                      const unsigned char* own) {
  __shared__ double sh[8];
  const int n = *G.n_active * D;
  double acc = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int t = k / D, i = k - t * D;
    const int A = G.act_list[t];
    double r = 0.0;
    if (!((fxr[t] >> i) & 1u)) {
      const double b = grav ? grav[(size_t)i * nsteps + step] : 0.0;
      r = -G.F[(size_t)A * D + i] + G.M[A] * (a1 * dU[k] - a2 * Vn[k] - a3 * An[k] - b);
    }
    R[k] = r;
    if (!own || own[t]) acc += r * r;
  }
  const double s = block_sum(acc, sh);
  if (threadIdx.x == 0) part_out[blockIdx.x] = s;
}

// ---- several slabs: ownership of the active nodes, band values of a compact vector through the force array
template <int D>
__global__ void __launch_bounds__(128) k_imp_own(MeshDev m, GridDev G, SlabDev sl, unsigned char* own) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *G.n_active) return;
  const double x = m.X[(size_t)G.act_list[t] * NS<D>::X + sl.axis];
  own[t] = (unsigned char)(x >= sl.own_lo && x < sl.own_hi);
}
template <int D>
__global__ void __launch_bounds__(256) k_band_put(GridDev G, const int* ids0, int n0, const int* ids1, int n1, const double* v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n0 + n1) return;
  const int A = i < n0 ? ids0[i] : ids1[i - n0];
  const int t = G.active[A] ? G.arank[A] : -1;
#pragma unroll
  for (int k = 0; k < D; k++) G.F[(size_t)A * D + k] = t >= 0 ? v[(size_t)t * D + k] : 0.0;
}
template <int D>
__global__ void __launch_bounds__(256) k_band_get(GridDev G, const int* ids0, int n0, const int* ids1, int n1, double* v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n0 + n1) return;
  const int A = i < n0 ? ids0[i] : ids1[i - n0];
  const int t = G.active[A] ? G.arank[A] : -1;
  if (t < 0) return;
#pragma unroll
  for (int k = 0; k < D; k++) v[(size_t)t * D + k] = G.F[(size_t)A * D + k];
}
__global__ void __launch_bounds__(256) k_diag_guard(const int* n_active, int D, double* diag) {
  const int n = *n_active * D;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
    if (diag[k] == 0.0) diag[k] = 1.0;  // a node that carries nothing on this slab and lies outside the exchanged band
}

// ---------------------------------------------------------------------------
// block-CSR pattern of the step: row t = active node, columns = ranks of the active nodes of its coupling set
__global__ void __launch_bounds__(256) k_csr_count(GridDev G, const int* cpl_ptr, const int* cpl_idx, ulonglong2* packed, int n_items) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_items) return;
  unsigned long long c = 0;
  if (t < *G.n_active) {
    const int A = G.act_list[t];
    for (int q = cpl_ptr[A]; q < cpl_ptr[A + 1]; q++) c += G.active[cpl_idx[q]] ? 1 : 0;
  }
  packed[t] = make_ulonglong2(c, 0ull);
}
__global__ void __launch_bounds__(256) k_csr_fill(GridDev G, const int* cpl_ptr, const int* cpl_idx, const int* row_ptr, int* cols) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *G.n_active) return;
  const int A = G.act_list[t];
  int o = row_ptr[t];
  for (int q = cpl_ptr[A]; q < cpl_ptr[A + 1]; q++) {
    const int B = cpl_idx[q];
    if (G.active[B]) cols[o++] = G.arank[B];  // cpl rows are sorted by node id, ranks are monotone in the id
  }
}

__device__ __forceinline__ int csr_find(const int* cols, int lo, int hi, int key) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int v = cols[mid];
    if (v == key) return mid;
    if (v < key) lo = mid + 1; else hi = mid;
  }
  return -1;
}

// Tangent values, one warp per particle (U-Newmark-beta.c:1646-1830 with compute_stiffness_density_Neo_Hookean,
// Neo-Hookean.c:89-141):  K_AB += V0 [ c0 g1_A (x) g1_B + G (g_B . b_n g_A) I + c1 g1_B (x) g1_A ],
// g = grad N at t_n, g1 = DF^-T g, b_n = F_n F_n^T, c0 = lambda J^2, c1 = G - lambda (J^2 - 1)/2, J = J_n1.
// EP: elastoplastic laws (Drucker-Prager, Matsuoka-Nakai, Von-Mises) and Hencky use compute_stiffness_elastoplastic__Constitutive__
// (Constitutive/Plasticity/Elastoplastic-Tangent-Matrix.c:42-160): spectral form with C_ep of the return mapping, the
// eigen-pairs of b_e_n1 and the eigenvalues of tau, plus the geometric term -tau (g1_B (x) g1_A); unsymmetric when the
// flow rule is not associated, hence BiCGStab below.  The law is read per particle (mixed clouds work).
template <int D, int W, bool EP>
__global__ void __launch_bounds__(128) k_assemble_nh(MeshDev m, PartDev P, GridDev G, const int* row_ptr, const int* cols, double* vals,
                                                     int* err, const __grid_constant__ MatTable mt, const AlmeDev al) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int NMAX = 32 * W, DD = D * D;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double* s_g = (double*)smem + (size_t)wib * NMAX * 3 * D;  // per neighbour: g | g1 | b_n g
  int* s_rank = (int*)((double*)smem + (size_t)wpb * NMAX * 3 * D) + wib * NMAX;
  const int np = P.ld;
  for (int p = blockIdx.x * wpb + wib; p < P.np; p += gridDim.x * wpb) {
    const int base = m.r2p[P.I0[p]];
    double xp[D], lam[D];
#pragma unroll
    for (int i = 0; i < D; i++) { xp[i] = P.x[i * np + p]; lam[i] = P.lam[i * np + p]; }
    const double beta = P.beta[p];
    const MetricB MB = metric_load<D>(al, P.ld, p);  // aLME: the particle's metric instead of beta
    uint32_t mk[W];
    int n = 0, off[W];
#pragma unroll
    for (int w = 0; w < W; w++) { mk[w] = P.mask[(size_t)w * np + p]; off[w] = n; n += __popc(mk[w]); }
    // pass 1: unnormalised weights of this lane's slots, sums over the warp
    double e_[W], l_[W][D], red[1 + D + DD];
#pragma unroll
    for (int i = 0; i < 1 + D + DD; i++) red[i] = 0.0;
    int node_[W];
#pragma unroll
    for (int w = 0; w < W; w++) {
      e_[w] = 0.0;
      node_[w] = -1;
      if ((mk[w] >> lane) & 1u) {
        const int node = m.r2i[base + w * 32 + lane];
        node_[w] = node;
        double XA[D], ll = 0.0, lx = 0.0;
        ldvec<D>(&m.X[(size_t)node * NS<D>::X], XA);
#pragma unroll
        for (int i = 0; i < D; i++) { l_[w][i] = xp[i] - XA[i]; ll += l_[w][i] * l_[w][i]; lx += l_[w][i] * lam[i]; }
        e_[w] = exp(-metric_q<D>(MB, beta, ll, l_[w]) + lx);
        red[0] += e_[w];
#pragma unroll
        for (int i = 0; i < D; i++) {
          red[1 + i] += e_[w] * l_[w][i];
#pragma unroll
          for (int j = 0; j < D; j++) red[1 + D + i * D + j] += e_[w] * l_[w][i] * l_[w][j];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 1 + D + DD; i++)
      for (int o = 16; o > 0; o >>= 1) red[i] += __shfl_xor_sync(0xffffffffu, red[i], o);
    const double Zi = 1.0 / red[0];
    double r[D], JJ[DD], Ji[DD];
#pragma unroll
    for (int i = 0; i < D; i++) r[i] = red[1 + i] * Zi;
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
      for (int j = 0; j < D; j++) JJ[i * D + j] = red[1 + D + i * D + j] * Zi - r[i] * r[j];
    inverse<D>(JJ, Ji);
    double DF[DD], DFi[DD], Fn[DD], bn[DD];
#pragma unroll
    for (int i = 0; i < DD; i++) { DF[i] = P.DF[(size_t)i * np + p]; Fn[i] = P.F_n[(size_t)i * np + p]; }
    inverse<D>(DF, DFi);
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
      for (int j = 0; j < D; j++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; k++) s += Fn[i * D + k] * Fn[j * D + k];
        bn[i * D + j] = s;
      }
    // pass 2: gradients of this lane's slots into the warp's compact neighbour table
#pragma unroll
    for (int w = 0; w < W; w++) {
      if (node_[w] < 0) continue;
      const int idx = off[w] + __popc(mk[w] & ((1u << lane) - 1u));
      const double pa = e_[w] * Zi;
      double g[D], g1[D], bg[D];
#pragma unroll
      for (int i = 0; i < D; i++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < D; j++) s += Ji[i * D + j] * l_[w][j];
        g[i] = -pa * s;
      }
#pragma unroll
      for (int i = 0; i < D; i++) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int j = 0; j < D; j++) { s1 += DFi[j * D + i] * g[j]; s2 += bn[i * D + j] * g[j]; }  // DF^-T g ; b_n g
        g1[i] = s1;
        bg[i] = s2;
      }
      s_rank[idx] = G.arank[node_[w]];
#pragma unroll
      for (int i = 0; i < D; i++) {
        s_g[(size_t)idx * 3 * D + i] = g[i];
        s_g[(size_t)idx * 3 * D + D + i] = g1[i];
        s_g[(size_t)idx * 3 * D + 2 * D + i] = bg[i];
      }
    }
    __syncwarp();
    const MatParams& mat = mt.m[P.matidx[p]];
    const double Gm = mat.E / (2 * (1 + mat.nu)), lm = mat.nu * mat.E / ((1 - mat.nu * 2) * (1 + mat.nu));
    const double J = P.J_n1[p], V0 = P.vol0[p];
    const double c0 = V0 * lm * J * J, c1 = V0 * (Gm - 0.5 * lm * (J * J - 1.0)), cg = V0 * Gm;
    if (EP && mat.type != NLPS_MAT_NEO_HOOKEAN_WRIGGERS) {
      double be[DD], ta[DD], cep[DD], lb[3], ev[9], lT[3], evT[9];
#pragma unroll
      for (int i = 0; i < DD; i++) {
        be[i] = P.be_n1[(size_t)i * np + p];
        ta[i] = P.stress[(size_t)i * np + p];
        cep[i] = P.cep[(size_t)i * np + p];
      }
      if (mat.type == NLPS_MAT_HENCKY) {
        // compute_stiffness_density_Hencky__Constitutive__ (Constitutive/Hyperelastic/Hencky.c:98-232): the same spectral
        // block with b = F_n1 F_n1^T and the constant elastic moduli AA in place of b_e_n1 and C_ep
        double F1[DD];
#pragma unroll
        for (int i = 0; i < DD; i++) F1[i] = P.F_n1[(size_t)i * np + p];
#pragma unroll
        for (int i = 0; i < D; i++)
#pragma unroll
          for (int j = 0; j < D; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; k++) s += F1[i * D + k] * F1[j * D + k];
            be[i * D + j] = s;
            cep[i * D + j] = (i == j) ? mat.lame + 2 * mat.G : mat.lame;
          }
      }
      if (D == 2) { dsyev2_dev(be[0], be[1], be[3], lb, ev); dsyev2_dev(ta[0], ta[1], ta[3], lT, evT); }
      else { jacobi3_dev(be, lb, ev); jacobi3_dev(ta, lT, evT); }
      for (int q = lane; q < n * n; q += 32) {
        const int a = q / n, b = q - a * n;
        const double* u = s_g + (size_t)a * 3 * D + D;  // g1 of node A (dN_alpha_n1)
        const double* v = s_g + (size_t)b * 3 * D + D;  // g1 of node B (dN_beta_n1)
        double Kd[DD], ue[D], ve[D];
#pragma unroll
        for (int i = 0; i < DD; i++) Kd[i] = 0.0;
#pragma unroll
        for (int A = 0; A < D; A++) {
          ue[A] = 0.0; ve[A] = 0.0;
#pragma unroll
          for (int i = 0; i < D; i++) { ue[A] += u[i] * ev[A + i * D]; ve[A] += v[i] * ev[A + i * D]; }
        }
#pragma unroll
        for (int A = 0; A < D; A++)
#pragma unroll
          for (int B = 0; B < D; B++) {
            const double C = cep[A * D + B];
            const bool geo = (A != B) && fabs(lb[B] - lb[A]) > 1E-14;
            const double rat = geo ? 0.5 * ((lT[B] - lT[A]) / (lb[B] - lb[A])) : 0.0;
#pragma unroll
            for (int i = 0; i < D; i++)
#pragma unroll
              for (int j = 0; j < D; j++) {
                Kd[i * D + j] += C * (ue[A] * ve[B]) * ev[A + i * D] * ev[B + j * D];
                if (geo)
                  Kd[i * D + j] += rat * (lb[B] * (ue[B] * ve[B]) * (ev[A + i * D] * ev[A + j * D]) +
                                          lb[A] * (ue[B] * ve[A]) * (ev[A + i * D] * ev[B + j * D]));
              }
          }
#pragma unroll
        for (int i = 0; i < D; i++)
#pragma unroll
          for (int j = 0; j < D; j++)
#pragma unroll
            for (int k = 0; k < D; k++) Kd[i * D + j] += -ta[i * D + k] * (v[k] * u[j]);
        const int row = s_rank[a];
        const int pos = csr_find(cols, row_ptr[row], row_ptr[row + 1], s_rank[b]);
        if (pos < 0) { latch_error(err, NLPS_ERR_CSR_PATTERN, P.orig[p]); continue; }
        double* dst = vals + (size_t)pos * DD;
#pragma unroll
        for (int i = 0; i < DD; i++) atomicAdd(&dst[i], V0 * Kd[i]);
      }
      __syncwarp();
      continue;
    }
    for (int q = lane; q < n * n; q += 32) {
      const int a = q / n, b = q - a * n;
      const double* ga = s_g + (size_t)a * 3 * D;
      const double* gb = s_g + (size_t)b * 3 * D;
      double len = 0.0;
#pragma unroll
      for (int i = 0; i < D; i++) len += gb[i] * ga[2 * D + i];
      const int row = s_rank[a];
      const int pos = csr_find(cols, row_ptr[row], row_ptr[row + 1], s_rank[b]);
      if (pos < 0) { latch_error(err, NLPS_ERR_CSR_PATTERN, P.orig[p]); continue; }
      double* dst = vals + (size_t)pos * DD;
#pragma unroll
      for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = 0; j < D; j++)
          atomicAdd(&dst[i * D + j], c0 * ga[D + i] * gb[D + j] + (i == j ? cg * len : 0.0) + c1 * ga[D + j] * gb[D + i]);
    }
    __syncwarp();
  }
}

// Neo-Hookean tangent, CELL-aggregated (default for clouds without elastoplastic laws): the particles of an occupied
// cell share their closest node, hence the 2-ring slot <-> node map, and most of their neighbours.  One block per cell:
// phase A, a warp per particle, fills a table [particle][slot] of (g, g1 = DF^-T g, b_n g) exactly as k_assemble_nh does;
// phase B, a thread per (slot a, slot b) pair of the union of the particles' lists, sums V0 K_AB over the cell's
// particles that hold both nodes and issues ONE set of d x d atomics per pair and cell instead of one per pair and
// particle (3D, 8 particles per cell: 2.8x fewer RED.E.ADD.F64, the unit that bounds the assembly).
template <int D, int W>
__global__ void __launch_bounds__(128) k_assemble_cell_nh(MeshDev m, PartDev P, GridDev G, const int* row_ptr, const int* cols,
                                                          double* vals, int* err, const __grid_constant__ MatTable mats,
                                                          const AlmeDev al) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int SL = 32 * W, DD = D * D, MP = 8, V = 3 * D;
  double* s_g = (double*)smem;                        // [MP][SL][V]: g | g1 | b_n g
  double* s_coef = s_g + (size_t)MP * SL * V;         // [MP][4]: c0, c1, cg (times V0)
  int* s_rank = (int*)(s_coef + MP * 4);              // [SL] active rank of the ring node of a slot
  uint32_t* s_mask = (uint32_t*)(s_rank + SL);        // [MP][W]
  int* s_ul = (int*)(s_mask + MP * W);                // [SL] slots in the union of the chunk's lists
  __shared__ int s_nu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int nocc = *G.n_occ, np = P.ld;
  for (int cell = blockIdx.x; cell < nocc; cell += gridDim.x) {
    const int4 mt = G.occ_meta[cell];
    const int t0 = mt.y, t1 = (cell + 1 < nocc) ? G.occ_meta[cell + 1].y : P.np;
    const int base = mt.z, len = mt.w & 511;
    __syncthreads();  // the previous cell's pairs are done with s_rank
    for (int k = threadIdx.x; k < SL; k += blockDim.x) s_rank[k] = (k < len) ? G.arank[m.r2i[base + k]] : -1;
    for (int tb = t0; tb < t1; tb += MP) {
      const int mc = min(MP, t1 - tb);
      __syncthreads();  // the previous chunk's pairs are done with the tables
      // ---- phase A: warp per particle (the weights and gradients of k_assemble_nh, dense by slot)
      for (int j = wib; j < mc; j += wpb) {
        const int p = G.plist[tb + j];
        double xp[D], lam[D];
#pragma unroll
        for (int i = 0; i < D; i++) { xp[i] = P.x[i * np + p]; lam[i] = P.lam[i * np + p]; }
        const double beta = P.beta[p];
        const MetricB MB = metric_load<D>(al, P.ld, p);  // aLME: the particle's metric instead of beta
        uint32_t mk[W];
#pragma unroll
        for (int w = 0; w < W; w++) mk[w] = P.mask[(size_t)w * np + p];
        double e_[W], l_[W][D], red[1 + D + DD];
#pragma unroll
        for (int i = 0; i < 1 + D + DD; i++) red[i] = 0.0;
#pragma unroll
        for (int w = 0; w < W; w++) {
          e_[w] = 0.0;
          if ((mk[w] >> lane) & 1u) {
            const int node = m.r2i[base + w * 32 + lane];
            double XA[D], ll = 0.0, lx = 0.0;
            ldvec<D>(&m.X[(size_t)node * NS<D>::X], XA);
#pragma unroll
            for (int i = 0; i < D; i++) { l_[w][i] = xp[i] - XA[i]; ll += l_[w][i] * l_[w][i]; lx += l_[w][i] * lam[i]; }
            e_[w] = exp(-metric_q<D>(MB, beta, ll, l_[w]) + lx);
            red[0] += e_[w];
#pragma unroll
            for (int i = 0; i < D; i++) {
              red[1 + i] += e_[w] * l_[w][i];
#pragma unroll
              for (int j2 = 0; j2 < D; j2++) red[1 + D + i * D + j2] += e_[w] * l_[w][i] * l_[w][j2];
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 1 + D + DD; i++)
          for (int o = 16; o > 0; o >>= 1) red[i] += __shfl_xor_sync(0xffffffffu, red[i], o);
        const double Zi = 1.0 / red[0];
        double r[D], JJ[DD], Ji[DD];
#pragma unroll
        for (int i = 0; i < D; i++) r[i] = red[1 + i] * Zi;
#pragma unroll
        for (int i = 0; i < D; i++)
#pragma unroll
          for (int j2 = 0; j2 < D; j2++) JJ[i * D + j2] = red[1 + D + i * D + j2] * Zi - r[i] * r[j2];
        inverse<D>(JJ, Ji);
        double DF[DD], DFi[DD], Fn[DD], bn[DD];
#pragma unroll
        for (int i = 0; i < DD; i++) { DF[i] = P.DF[(size_t)i * np + p]; Fn[i] = P.F_n[(size_t)i * np + p]; }
        inverse<D>(DF, DFi);
#pragma unroll
        for (int i = 0; i < D; i++)
#pragma unroll
          for (int j2 = 0; j2 < D; j2++) {
            double s_ = 0.0;
#pragma unroll
            for (int k = 0; k < D; k++) s_ += Fn[i * D + k] * Fn[j2 * D + k];
            bn[i * D + j2] = s_;
          }
#pragma unroll
        for (int w = 0; w < W; w++) {
          if (!((mk[w] >> lane) & 1u)) continue;
          const double pa = e_[w] * Zi;
          double g[D];
#pragma unroll
          for (int i = 0; i < D; i++) {
            double s_ = 0.0;
#pragma unroll
            for (int j2 = 0; j2 < D; j2++) s_ += Ji[i * D + j2] * l_[w][j2];
            g[i] = -pa * s_;
          }
          double* dst = s_g + ((size_t)j * SL + w * 32 + lane) * V;
#pragma unroll
          for (int i = 0; i < D; i++) {
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int j2 = 0; j2 < D; j2++) { s1 += DFi[j2 * D + i] * g[j2]; s2 += bn[i * D + j2] * g[j2]; }  // DF^-T g ; b_n g
            dst[i] = g[i];
            dst[D + i] = s1;
            dst[2 * D + i] = s2;
          }
        }
        if (lane == 0) {
          const MatParams& mat = mats.m[P.matidx[p]];
          const double Gm = mat.E / (2 * (1 + mat.nu)), lm = mat.nu * mat.E / ((1 - mat.nu * 2) * (1 + mat.nu));
          const double J = P.J_n1[p], V0 = P.vol0[p];
          s_coef[j * 4 + 0] = V0 * lm * J * J;
          s_coef[j * 4 + 1] = V0 * (Gm - 0.5 * lm * (J * J - 1.0));
          s_coef[j * 4 + 2] = V0 * Gm;
        }
        if (lane < W) s_mask[j * W + lane] = P.mask[(size_t)lane * np + p];
      }
      __syncthreads();
      // ---- union of the chunk's neighbour lists (ascending slots)
      if (wib == 0) {
        int cnt = 0;
#pragma unroll
        for (int w = 0; w < W; w++) {
          uint32_t u = 0u;
          for (int j = 0; j < mc; j++) u |= s_mask[j * W + w];
          if ((u >> lane) & 1u) s_ul[cnt + __popc(u & ((1u << lane) - 1u))] = w * 32 + lane;
          cnt += __popc(u);
        }
        if (lane == 0) s_nu = cnt;
      }
      __syncthreads();
      // ---- phase B: thread per pair of union slots, summed over the particles that hold both
      const int nu = s_nu;
      for (int q = threadIdx.x; q < nu * nu; q += blockDim.x) {
        const int qa = q / nu;
        const int a = s_ul[qa], b = s_ul[q - qa * nu];
        const int wa = a >> 5, wb = b >> 5;
        const uint32_t ba = 1u << (a & 31), bb = 1u << (b & 31);
        double K[DD];
#pragma unroll
        for (int i = 0; i < DD; i++) K[i] = 0.0;
        bool any = false;
        for (int j = 0; j < mc; j++) {
          if (!(s_mask[j * W + wa] & ba) || !(s_mask[j * W + wb] & bb)) continue;
          any = true;
          const double* ga = s_g + ((size_t)j * SL + a) * V;
          const double* gb = s_g + ((size_t)j * SL + b) * V;
          const double c0 = s_coef[j * 4 + 0], c1 = s_coef[j * 4 + 1], cg = s_coef[j * 4 + 2];
          double ln = 0.0;
#pragma unroll
          for (int i = 0; i < D; i++) ln += gb[i] * ga[2 * D + i];
#pragma unroll
          for (int i = 0; i < D; i++)
#pragma unroll
            for (int j2 = 0; j2 < D; j2++)
              K[i * D + j2] += c0 * ga[D + i] * gb[D + j2] + (i == j2 ? cg * ln : 0.0) + c1 * ga[D + j2] * gb[D + i];
        }
        if (!any) continue;  // nobody holds both nodes: the pair may not even be in the pattern
        const int row = s_rank[a];
        const int pos = csr_find(cols, row_ptr[row], row_ptr[row + 1], s_rank[b]);
        if (pos < 0) { latch_error(err, NLPS_ERR_CSR_PATTERN, P.orig[G.plist[tb]]); continue; }
        double* dst = vals + (size_t)pos * DD;
#pragma unroll
        for (int i = 0; i < DD; i++) atomicAdd(&dst[i], K[i]);
      }
    }
  }
}

// Jacobi preconditioner: diagonal of (K + alpha_1 M) with unit rows on restricted dofs
template <int D>
__global__ void __launch_bounds__(128) k_bsr_diag(GridDev G, const int* row_ptr, const int* cols, const double* vals,
                                                  const unsigned char* fxr, double a1, double* diag, const unsigned char* own) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *G.n_active) return;
  const int pos = csr_find(cols, row_ptr[t], row_ptr[t + 1], t);
  const double M = G.M[G.act_list[t]];
  const double ow = (!own || own[t]) ? 1.0 : 0.0;  // the mass term and the unit rows are the owner's share of the band sum
#pragma unroll
  for (int i = 0; i < D; i++) {
    const double k = pos >= 0 ? vals[(size_t)pos * D * D + i * D + i] : 0.0;
    diag[(size_t)t * D + i] = ((fxr[t] >> i) & 1u) ? ow : k + ow * a1 * M;
  }
}

// y = (K + alpha_1 M) x with the Dirichlet rows and columns replaced by the identity (MatZeroRowsColumnsIS,
// U-Newmark-beta.c:1828), one warp per block row, lanes over the scalar entries of the row (coalesced);
// part_out[block] = partial sum of x.y
// part_w[block] = partial sum of w.y (w = x for CG), part_yy[block] (optional) of y.y
template <int D>
__global__ void __launch_bounds__(256) k_bsr_spmv(GridDev G, const int* row_ptr, const int* cols, const double* vals,
                                                  const unsigned char* fxr, double a1, const double* x, double* y,
                                                  const double* w, double* part_out, double* part_yy,
                                                  const unsigned char* own = nullptr) {
  __shared__ double sh[8];
  constexpr int DD = D * D;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int nact = *G.n_active;
  double dot = 0.0, dyy = 0.0;
  // lane <-> (block of the row, row i of the block): D contiguous doubles of K and the D values of x per lane and
  // block (a third of the instructions of a lane-per-scalar-entry mapping: the kernel is bound by issue slots before HBM)
  constexpr int BPW = 32 / D;               // blocks per warp sweep (10 in 3D: lanes 30, 31 idle)
  const int sub = lane % D, lb = lane / D;
  for (int t = blockIdx.x * wpb + (threadIdx.x >> 5); t < nact; t += gridDim.x * wpb) {
    const int q0 = row_ptr[t], nb = row_ptr[t + 1] - q0;
    const double* v = vals + (size_t)q0 * DD + sub * D;
    double acc = 0.0;
    if (lb < BPW) {
      for (int blk = lb; blk < nb; blk += BPW) {
        const int c = cols[q0 + blk];
        const unsigned fxc = fxr[c];
        const double* kv = v + (size_t)blk * DD;
        const double* xv = x + (size_t)c * D;
#pragma unroll
        for (int j = 0; j < D; j++) acc += kv[j] * (((fxc >> j) & 1u) ? 0.0 : xv[j]);
      }
    }
    // sum over the lanes of the same block row (lane stride D); BPW need not be a power of two
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      if (o >= BPW) continue;
      const double other = __shfl_down_sync(0xffffffffu, acc, o * D);
      if (lb < o && lb + o < ((2 * o < BPW) ? 2 * o : BPW)) acc += other;
    }
    if (lane < D) {
      const int i = lane;
      const double M = G.M[G.act_list[t]];
      const double xi = x[(size_t)t * D + i];
      // several slabs: y is this slab's share of the band sum -- the mass term and the unit rows belong to the owner; the
      // partial of w.y then runs over ALL local rows (w^T K w = sum over the slabs of w^T K_slab w)
      const double ow = (!own || own[t]) ? 1.0 : 0.0;
      const double yi = ((fxr[t] >> i) & 1u) ? ow * xi : acc + ow * a1 * M * xi;
      y[(size_t)t * D + i] = yi;
      dot += w[(size_t)t * D + i] * yi;
      dyy += yi * yi;
    }
  }
  const double s = block_sum(dot, sh);
  if (threadIdx.x == 0) part_out[blockIdx.x] = s;
  if (part_yy) {
    const double s2 = block_sum(dyy, sh);
    if (threadIdx.x == 0) part_yy[blockIdx.x] = s2;
  }
}

// PCG vector kernels; every scalar is re-summed from the partial arrays in a fixed order by every block
// x = 0, r = b, z = r / diag, p = z; partials of r.z -> part_rz, of r.r -> part_rr
__global__ void __launch_bounds__(256) k_pcg_init(const int* n_active, int D, const double* b, const double* diag, double* x, double* r,
                                                  double* z, double* p, double* part_rz, double* part_rr,
                                                  const unsigned char* own = nullptr) {
  __shared__ double sh[8];
  const int n = *n_active * D;
  double a = 0.0, c = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const double rv = b[k], zv = rv / diag[k];
    x[k] = 0.0; r[k] = rv; z[k] = zv; p[k] = zv;
    if (own && !own[k / D]) continue;  // dot products count every node once: on the slab that owns it
    a += rv * zv;
    c += rv * rv;
  }
  const double s1 = block_sum(a, sh), s2 = block_sum(c, sh);
  if (threadIdx.x == 0) { part_rz[blockIdx.x] = s1; part_rr[blockIdx.x] = s2; }
}
__global__ void __launch_bounds__(256) k_pcg_update1(const int* n_active, int D, const double* part_pAp, const double* part_rz_cur,
                                                     const double* p, const double* Ap, const double* diag, double* x, double* r,
                                                     double* z, double* part_rz_new, double* part_rr,
                                                     const unsigned char* own = nullptr) {
  __shared__ double sh[8];
  const int n = *n_active * D;
  const double pAp = sum_partials(part_pAp, IMP_NSPMV), rz = sum_partials(part_rz_cur);
  const double alpha = (pAp != 0.0) ? rz / pAp : 0.0;
  double a = 0.0, c = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    x[k] += alpha * p[k];
    const double rv = r[k] - alpha * Ap[k], zv = rv / diag[k];
    r[k] = rv; z[k] = zv;
    if (own && !own[k / D]) continue;
    a += rv * zv;
    c += rv * rv;
  }
  const double s1 = block_sum(a, sh), s2 = block_sum(c, sh);
  if (threadIdx.x == 0) { part_rz_new[blockIdx.x] = s1; part_rr[blockIdx.x] = s2; }
}
__global__ void __launch_bounds__(256) k_pcg_update2(const int* n_active, int D, const double* part_rz_cur, const double* part_rz_new,
                                                     const double* z, double* p) {
  const int n = *n_active * D;
  const double rz = sum_partials(part_rz_cur), rzn = sum_partials(part_rz_new);
  const double beta = (rz != 0.0) ? rzn / rz : 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) p[k] = z[k] + beta * p[k];
}
// Jacobi-preconditioned BiCGStab for the unsymmetric elastoplastic tangents; scalars as in the PCG: every block
// re-sums the per-block partials of the previous kernels in a fixed order (q = parity of the iteration)
__device__ __forceinline__ double safe_div(double a, double b) { return b != 0.0 ? a / b : 0.0; }
struct BiParts {  // [2] = ping-pong by iteration parity
  double *rho[2], *rv[2], *ts[2], *tt[2], *rr;
};
__global__ void __launch_bounds__(256) k_bi_init(const int* n_active, int D, const double* b, double* x, double* r, double* rhat,
                                                 double* part_rho0, double* part_rr) {
  __shared__ double sh[8];
  const int n = *n_active * D;
  double a = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const double bv = b[k];
    x[k] = 0.0; r[k] = bv; rhat[k] = bv;
    a += bv * bv;
  }
  const double s = block_sum(a, sh);
  if (threadIdx.x == 0) { part_rho0[blockIdx.x] = s; part_rr[blockIdx.x] = s; }
}
__global__ void __launch_bounds__(256) k_bi_p(const int* n_active, int D, BiParts P_, int it, const double* r, const double* v,
                                              const double* diag, double* p, double* y) {
  const int n = *n_active * D, q = it & 1, pq = q ^ 1;
  double beta = 0.0, omega_prev = 0.0;
  if (it > 0) {
    const double rho = sum_partials(P_.rho[q]), rho_prev = sum_partials(P_.rho[pq]);
    const double alpha_prev = safe_div(rho_prev, sum_partials(P_.rv[pq], IMP_NSPMV));
    omega_prev = safe_div(sum_partials(P_.ts[pq], IMP_NSPMV), sum_partials(P_.tt[pq], IMP_NSPMV));
    beta = safe_div(rho, rho_prev) * safe_div(alpha_prev, omega_prev);
  }
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const double pv = (it > 0) ? r[k] + beta * (p[k] - omega_prev * v[k]) : r[k];
    p[k] = pv;
    y[k] = pv / diag[k];
  }
}
__global__ void __launch_bounds__(256) k_bi_s(const int* n_active, int D, BiParts P_, int it, const double* r, const double* v,
                                              const double* diag, double* s, double* z) {
  const int n = *n_active * D, q = it & 1;
  const double alpha = safe_div(sum_partials(P_.rho[q]), sum_partials(P_.rv[q], IMP_NSPMV));
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const double sv = r[k] - alpha * v[k];
    s[k] = sv;
    z[k] = sv / diag[k];
  }
}
__global__ void __launch_bounds__(256) k_bi_x(const int* n_active, int D, BiParts P_, int it, const double* y, const double* z,
                                              const double* s, const double* t, const double* rhat, double* x, double* r) {
  __shared__ double sh[8];
  const int n = *n_active * D, q = it & 1;
  const double alpha = safe_div(sum_partials(P_.rho[q]), sum_partials(P_.rv[q], IMP_NSPMV));
  const double omega = safe_div(sum_partials(P_.ts[q], IMP_NSPMV), sum_partials(P_.tt[q], IMP_NSPMV));
  double a = 0.0, c = 0.0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    x[k] += alpha * y[k] + omega * z[k];
    const double rv = s[k] - omega * t[k];
    r[k] = rv;
    a += rhat[k] * rv;
    c += rv * rv;
  }
  const double s1 = block_sum(a, sh), s2 = block_sum(c, sh);
  if (threadIdx.x == 0) { P_.rho[q ^ 1][blockIdx.x] = s1; P_.rr[blockIdx.x] = s2; }
}
// out = a + s * b ; out = -a
__global__ void __launch_bounds__(256) k_vec_axpy(const int* n_active, int D, const double* a, double s, const double* b, double* out) {
  const int n = *n_active * D;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) out[k] = a[k] + s * b[k];
}
__global__ void __launch_bounds__(256) k_vec_neg(const int* n_active, int D, const double* a, double* out) {
  const int n = *n_active * D;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) out[k] = -a[k];
}

// G3 + K4: nodal kinetic increments (U-Newmark-beta.c:1859-1906), FLIP update of the particles (:1993-2072,
// alpha_blend = 1) and rho = m / (V0 J) (:1929-1932); thread per particle
template <int D, int W>
__global__ void __launch_bounds__(128) k_g2p_implicit(MeshDev m, PartDev P, GridDev G, double a1, double a2, double a3, double a4,
                                                      double a5, double a6, const double* dU, const double* Vn, const double* An,
                                                      const AlmeDev al) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.np) return;
  const int np = P.ld, base = m.r2p[P.I0[p]];
  double xp[D], lam[D], sU[D], sV[D], sA[D], Z = 0.0;
#pragma unroll
  for (int i = 0; i < D; i++) { xp[i] = P.x[i * np + p]; lam[i] = P.lam[i * np + p]; sU[i] = sV[i] = sA[i] = 0.0; }
  const double beta = P.beta[p];
  const MetricB MB = metric_load<D>(al, P.ld, p);  // aLME: the particle's metric instead of beta
#pragma unroll
  for (int w = 0; w < W; w++) {
    uint32_t mm = P.mask[(size_t)w * np + p];
    while (mm) {
      const int k = w * 32 + __ffs(mm) - 1;
      mm &= mm - 1;
      const int node = m.r2i[base + k];
      double XA[D], ll = 0.0, lx = 0.0;
      ldvec<D>(&m.X[(size_t)node * NS<D>::X], XA);
      double lv[D];
#pragma unroll
      for (int i = 0; i < D; i++) { lv[i] = xp[i] - XA[i]; ll += lv[i] * lv[i]; lx += lv[i] * lam[i]; }
      const double e = exp(-metric_q<D>(MB, beta, ll, lv) + lx);
      const size_t t = (size_t)G.arank[node];
      Z += e;
#pragma unroll
      for (int i = 0; i < D; i++) {
        const double u = dU[t * D + i], v = Vn[t * D + i], a = An[t * D + i];
        sU[i] += e * u;
        sV[i] += e * (a4 * u + (a5 - 1.0) * v + a6 * a);
        sA[i] += e * (a1 * u - a2 * v - (a3 + 1.0) * a);
      }
    }
  }
  const double Zi = 1.0 / Z;
#pragma unroll
  for (int i = 0; i < D; i++) {
    const double du = sU[i] * Zi;
    P.acc[i * np + p] += sA[i] * Zi;
    P.vel[i * np + p] += sV[i] * Zi;
    P.dis[i * np + p] += du;
    P.x[i * np + p] = xp[i] + du;
    P.ddis[i * np + p] = du;
  }
  P.rho[p] = P.mass[p] / (P.vol0[p] * P.J_n1[p]);
}

// compact vector (by active rank) -> full-grid array (nn x D), zero on inactive nodes
__global__ void k_imp_export(GridDev G, int nn, int D, const double* v, double* out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)nn * D) return;
  const int A = (int)(i / D), k = (int)(i % D);
  out[i] = G.active[A] ? v[(size_t)G.arank[A] * D + k] : 0.0;
}
__global__ void k_imp_import(GridDev G, int nn, int D, const double* in, double* v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)nn * D) return;
  const int A = (int)(i / D), k = (int)(i % D);
  if (G.active[A]) v[(size_t)G.arank[A] * D + k] = in[i];
}
__global__ void k_csr_cols_to_nodes(GridDev G, const int* cols, int n, int* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = G.act_list[cols[i]];
}

// ---------------------------------------------------------------------------
// host side
static int imp_setup(nlps_engine* e, const nlps_newmark* prm) {
  implicit_free(e);
  if (e->slab_on && e->uniform_mat != NLPS_MAT_NEO_HOOKEAN_WRIGGERS) {
    fprintf(stderr, "nlps_b200_newmark_setup: several slabs run the symmetric (Neo-Hookean) operator only; the elastoplastic "
                    "tangents (Jacobi-BiCGStab) run on a single slab\n");
    return 1;
  }
  if (!prm->quasi_static && (!(prm->beta > 0.0) || !(prm->gamma > 0.0))) {  // a1 = 1/(beta dt^2): explicit central differences are U_Verlet's job
    fprintf(stderr, "nlps_b200_newmark_setup: beta and gamma must be positive (beta=%g gamma=%g)\n", prm->beta, prm->gamma);
    return 1;
  }
  e->imp = new ImplicitCtx();
  ImplicitCtx* c = e->imp;
  c->prm = *prm;
  c->plastic = e->uniform_mat != NLPS_MAT_NEO_HOOKEAN_WRIGGERS;
  if (c->plastic) e->solver.compute_c_ep = 1;  // the elastoplastic tangent reads Phi.C_ep (Constitutive.c:330-355)
  if (c->prm.pcg_rtol <= 0.0) c->prm.pcg_rtol = 1e-8;
  if (c->prm.pcg_max_iter <= 0) c->prm.pcg_max_iter = 10000;
  const double dt = e->dt, b = prm->beta, g = prm->gamma;
  c->a1 = 1 / (b * dt * dt);  // __compute_Newmark_parameters (U-Newmark-beta.c:497-514)
  c->a2 = 1 / (b * dt);
  c->a3 = (1 - 2 * b) / (2 * b);
  c->a4 = g / (b * dt);
  c->a5 = 1 - g / b;
  c->a6 = (1 - g / (2 * b)) * dt;
  if (prm->quasi_static) {  // U_Static (U-Static.c): no inertia; v_n = a_n = 0 below make every other term vanish
    c->a1 = c->a2 = c->a4 = c->a6 = 0.0;
    c->a3 = -1.0;  // G2P: dA = -(a3 + 1) a_n
    c->a5 = 1.0;   // G2P: dV = (a5 - 1) v_n
  }
  const int nn = e->nn, D = e->D, xs = (D == 2) ? 2 : 4;
  // ---- static coupling adjacency on the host: B couples with A when one cell's 2-ring holds both and they are
  // closer than two support radii (a particle can have both in its list); sorted by node id
  std::vector<int> r2p(nn + 1), tp(nn + 1);
  CUDA_OK(cudaMemcpy(r2p.data(), e->mesh.r2p, sizeof(int) * (nn + 1), cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(tp.data(), e->mesh.r2tp, sizeof(int) * (nn + 1), cudaMemcpyDeviceToHost));
  std::vector<int> r2i(r2p[nn]), ti(tp[nn]);
  std::vector<double> X((size_t)nn * xs), h(nn);
  CUDA_OK(cudaMemcpy(r2i.data(), e->mesh.r2i, sizeof(int) * r2i.size(), cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(ti.data(), e->mesh.r2ti, sizeof(int) * ti.size(), cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(X.data(), e->mesh.X, sizeof(double) * X.size(), cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(h.data(), e->mesh.h_avg, sizeof(double) * nn, cudaMemcpyDeviceToHost));
  double hmax = 0.0;
  for (int i = 0; i < nn; i++) hmax = std::max(hmax, h[i]);
  const double Ra = hmax * sqrt(e->neg_log_tol / e->solver.gamma_lme), reach2 = 4.0 * Ra * Ra * (1.0 + 1e-9);
  std::vector<int> cp(nn + 1, 0);
  std::vector<std::vector<int>> rows(nn);
#pragma omp parallel
  {
    std::vector<int> stamp(nn, -1), tmp;
#pragma omp for schedule(dynamic, 256)
    for (int A = 0; A < nn; A++) {
      tmp.clear();
      for (int q = tp[A]; q < tp[A + 1]; q++) {
        const int C = ti[q];
        for (int s_ = r2p[C]; s_ < r2p[C + 1]; s_++) {
          const int B = r2i[s_];
          if (stamp[B] == A) continue;
          stamp[B] = A;
          double d2 = 0.0;
          for (int k = 0; k < D; k++) { const double dd = X[(size_t)A * xs + k] - X[(size_t)B * xs + k]; d2 += dd * dd; }
          if (d2 <= reach2) tmp.push_back(B);
        }
      }
      std::sort(tmp.begin(), tmp.end());
      rows[A] = tmp;
    }
  }
  size_t tot = 0;
  int maxrow = 0;
  for (int A = 0; A < nn; A++) { cp[A] = (int)tot; tot += rows[A].size(); maxrow = std::max(maxrow, (int)rows[A].size()); }
  if (tot > 0x7fffffffull) { fprintf(stderr, "nlps_b200 (implicit): coupling adjacency too large\n"); return 1; }
  cp[nn] = (int)tot;
  std::vector<int> ci(tot);
  for (int A = 0; A < nn; A++) std::copy(rows[A].begin(), rows[A].end(), ci.begin() + cp[A]);
  rows.clear();
  c->cpl_total = tot;
  const size_t nv = (size_t)e->max_act * D;
  c->cap_blocks = std::min(tot, (size_t)e->max_act * maxrow);
  if (imp_alloc(e, &c->cpl_ptr, (size_t)nn + 1) || imp_alloc(e, &c->cpl_idx, tot) || imp_alloc(e, &c->row_ptr, (size_t)e->max_act + 2) ||
      imp_alloc(e, &c->cols, c->cap_blocks) || imp_alloc(e, &c->vals, c->cap_blocks * D * D) ||
      imp_alloc(e, &c->dummy_a, (size_t)e->max_act + 2) || imp_alloc(e, &c->dummy_b, (size_t)e->max_act + 2) ||
      imp_alloc(e, &c->tops, 4) || imp_alloc(e, &c->packed, (size_t)e->max_act + 2) ||
      imp_alloc(e, &c->scan_blk, (size_t)nblk((size_t)e->max_act + 1, SCAN_ITEMS) + 1) || imp_alloc(e, &c->Vn, nv) ||
      imp_alloc(e, &c->An, nv) || imp_alloc(e, &c->dU, nv) || imp_alloc(e, &c->R, nv) || imp_alloc(e, &c->delta, nv) ||
      imp_alloc(e, &c->trial, nv) || imp_alloc(e, &c->Rt, nv) || imp_alloc(e, &c->r, nv) || imp_alloc(e, &c->z, nv) ||
      imp_alloc(e, &c->p, nv) || imp_alloc(e, &c->Ap, nv) || imp_alloc(e, &c->diag, nv) || imp_alloc(e, &c->fx, e->max_act) ||
      imp_alloc(e, &c->part, 4 * IMP_NPART + IMP_NSPMV))
    return 1;
  if (e->slab_on && (imp_alloc(e, &c->own, e->max_act) || imp_alloc(e, &c->pair, 4 * IMP_NPART) ||
                     imp_alloc(e, &c->ar_tab, (size_t)(e->world + 1) * IMP_NSPMV)))
    return 1;
  if (c->plastic && (imp_alloc(e, &c->bv, nv) || imp_alloc(e, &c->bs, nv) || imp_alloc(e, &c->bt, nv) || imp_alloc(e, &c->by, nv) ||
                     imp_alloc(e, &c->brh, nv) || imp_alloc(e, &c->part2, 3 * IMP_NPART + 6 * IMP_NSPMV)))
    return 1;
  CUDA_OK(cudaMemcpyAsync(c->cpl_ptr, cp.data(), sizeof(int) * (nn + 1), cudaMemcpyHostToDevice, e->stream));
  CUDA_OK(cudaMemcpyAsync(c->cpl_idx, ci.data(), sizeof(int) * tot, cudaMemcpyHostToDevice, e->stream));
  CUDA_OK(cudaMallocHost(&c->h_part, sizeof(double) * IMP_NPART));
  CUDA_OK(cudaStreamSynchronize(e->stream));
  return 0;
}

// several slabs: element-wise sum of a partial array over the slabs (afterwards identical everywhere), and the band sums
// of a compact nodal vector (carried through G.F and the force exchange of the explicit scheme)
static int imp_allreduce(nlps_engine* e, double* d_vec, int n) {
  if (!e->slab_on) return 0;
  if (comm_allreduce_sum(e->comm, e->rank, e->world, d_vec, n, e->imp->ar_tab, e->stream)) { e->host_fail = 1; return 1; }
  e->launches++;
  return 0;
}
template <int D>
static int imp_band_sum(nlps_engine* e, double* v) {
  if (!e->slab_on) return 0;
  const int n0 = e->side[0].peer >= 0 ? e->side[0].n : 0, n1 = e->side[1].peer >= 0 ? e->side[1].n : 0;
  if (n0 + n1 > 0) k_band_put<D><<<nblk(n0 + n1, 256), 256, 0, e->stream>>>(e->G, e->side[0].ids, n0, e->side[1].ids, n1, v);
  const int rc = halo_exchange<D>(e, 2);
  if (n0 + n1 > 0) k_band_get<D><<<nblk(n0 + n1, 256), 256, 0, e->stream>>>(e->G, e->side[0].ids, n0, e->side[1].ids, n1, v);
  e->launches += 2;
  return rc;
}
// a device error on one slab must stop every slab (the others would wait in the next collective)
static int imp_poll(nlps_engine* e) {
  const int bad = poll_error(e);
  if (!e->slab_on) return bad;
  int all_ok = 1;
  if (comm_all_ok(e->comm, e->rank, e->world, !bad, e->mig_cnt + 12, e->stream, &all_ok)) return 1;
  if (!all_ok && !bad) { e->host_fail = 1; if (!e->last_code) e->last_code = NLPS_ERR_CUDA; }
  return bad || !all_ok;
}

static double imp_norm_at(nlps_engine* e, double* d_part) {  // sqrt of the sum of a partial array (over all slabs)
  ImplicitCtx* c = e->imp;
  imp_allreduce(e, d_part, IMP_NPART);
  cudaMemcpyAsync(c->h_part, d_part, sizeof(double) * IMP_NPART, cudaMemcpyDeviceToHost, e->stream);
  cudaStreamSynchronize(e->stream);
  double s = 0.0;
  for (int i = 0; i < IMP_NPART; i++) s += c->h_part[i];
  return sqrt(s);
}
static double imp_norm(nlps_engine* e, int slot) {  // sqrt of the sum of a partial array
  ImplicitCtx* c = e->imp;
  if (e->slab_on) return imp_norm_at(e, c->part + (size_t)slot * IMP_NPART);
  cudaMemcpyAsync(c->h_part, c->part + (size_t)slot * IMP_NPART, sizeof(double) * IMP_NPART, cudaMemcpyDeviceToHost, e->stream);
  cudaStreamSynchronize(e->stream);
  double s = 0.0;
  for (int i = 0; i < IMP_NPART; i++) s += c->h_part[i];
  return sqrt(s);
}

// K1 of the implicit step: search, lumped mass, nodal v_n / a_n, restricted dofs, block-CSR pattern, initial guess
template <int D>
static int imp_begin_t(nlps_engine* e, int step) {
  ImplicitCtx* c = e->imp;
  GridDev& G = e->G;
  const int nb128 = nblk(e->max_act, 128);
  stage_search_t<D>(e, step, 1, 0, e->P.vel, 0);
  launch_grid_disp<D, 1>(e, step);
  halo_exchange<D>(e, 1);  // (several slabs: band sums of mass and momentum, as in the explicit scheme)
  k_imp_nodal<D><<<nb128, 128, 0, e->stream>>>(G, e->bc, step, 1, c->Vn, c->fx);
  stage_search_t<D>(e, step, 1, 0, e->P.acc, 1);
  launch_grid_disp<D, 1>(e, step);
  halo_exchange<D>(e, 1);
  k_imp_nodal<D><<<nb128, 128, 0, e->stream>>>(G, e->bc, step, 0, c->An, c->fx);
  if (e->slab_on) k_imp_own<D><<<nb128, 128, 0, e->stream>>>(e->mesh, G, slab_dev(e), c->own);
  if (c->prm.quasi_static) {  // the projections above are kept for the lumped mass and the restricted-DOF flags
    cudaMemsetAsync(c->Vn, 0, sizeof(double) * (size_t)e->max_act * D, e->stream);
    cudaMemsetAsync(c->An, 0, sizeof(double) * (size_t)e->max_act * D, e->stream);
  }
  // pattern
  const int items = e->max_act + 1, nb = nblk(items, SCAN_ITEMS);
  k_csr_count<<<nblk(items, 256), 256, 0, e->stream>>>(G, c->cpl_ptr, c->cpl_idx, c->packed, items);
  k_scan_reduce<<<nb, 256, 0, e->stream>>>(c->packed, c->scan_blk, items);
  k_scan_tops<<<1, 1024, 0, e->stream>>>(c->scan_blk, nb, c->tops, c->tops + 1, c->tops + 2);
  k_scan_apply<<<nb, 256, 0, e->stream>>>(c->packed, c->scan_blk, c->row_ptr, c->dummy_a, c->dummy_b, items);
  k_csr_fill<<<nblk(e->max_act, 256), 256, 0, e->stream>>>(G, c->cpl_ptr, c->cpl_idx, c->row_ptr, c->cols);
  k_imp_guess<D><<<nb128, 128, 0, e->stream>>>(G, e->bc, step, e->dt, c->prm.use_explicit_trial, c->Vn, c->An, c->dU);
  e->launches += 9;
  return 0;
}

// residual at the nodal increment `dU` (compact): leaves DF, F_n1, J_n1, stress of the particles at that state
template <int D>
static void imp_residual_t(nlps_engine* e, int step, const double* dU, double* R, int part_slot) {
  ImplicitCtx* c = e->imp;
  const int nb128 = nblk(e->max_act, 128);
  k_imp_set_dU<D><<<nb128, 128, 0, e->stream>>>(e->G, dU);
  e->implicit_on = 1;
  stage_kin_stress_t<D>(e, step);
  e->implicit_on = 0;
  launch_grid_acc<D, 1>(e, step);
  halo_exchange<D>(e, 2);  // (several slabs: band sums of the internal forces)
  k_imp_residual<D><<<IMP_NPART, 256, 0, e->stream>>>(e->G, e->grav, e->solver.num_steps, step, c->a1, c->a2, c->a3, dU, c->Vn, c->An,
                                                     c->fx, R, c->part + (size_t)part_slot * IMP_NPART, c->own);
  e->launches += 2;
  c->residual_evals++;
}

template <int D>
static int imp_assemble_t(nlps_engine* e) {
  ImplicitCtx* c = e->imp;
  cudaMemsetAsync(c->vals, 0, sizeof(double) * c->cap_blocks * D * D, e->stream);
  const int wpb = 4, threads = 32 * wpb;
  auto launch = [&](auto kfn, int W) {
    const size_t smem = (size_t)wpb * 32 * W * (3 * D * sizeof(double) + sizeof(int));
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = std::max(1, std::min(nblk(e->np, wpb), e->sm_count * 8));
    kfn<<<grid, threads, smem, e->stream>>>(e->mesh, e->P, e->G, c->row_ptr, c->cols, c->vals, e->err, e->mat, e->alme);
  };
  auto launch_cell = [&](auto kfn, int W) {
    const int SL = 32 * W, MP = 8;
    const size_t smem = sizeof(double) * ((size_t)MP * SL * 3 * D + MP * 4) + sizeof(int) * ((size_t)SL + MP * W + SL);
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int per_sm = std::max(1, std::min(8, (int)((size_t)e->max_smem_optin / (smem + 1024))));
    const int grid = std::max(1, std::min(std::max(e->max_occ, 1), e->sm_count * per_sm));
    kfn<<<grid, threads, smem, e->stream>>>(e->mesh, e->P, e->G, c->row_ptr, c->cols, c->vals, e->err, e->mat, e->alme);
  };
  static const bool cell_asm = !(getenv("NLPS_ASM_CELL") && atoi(getenv("NLPS_ASM_CELL")) == 0);
  if (!c->plastic && cell_asm) {
    if constexpr (D == 2) { if (e->W == 1) launch_cell(k_assemble_cell_nh<2, 1>, 1); else launch_cell(k_assemble_cell_nh<2, 2>, 2); }
    else { if (e->W == 4) launch_cell(k_assemble_cell_nh<3, 4>, 4); else launch_cell(k_assemble_cell_nh<3, 8>, 8); }
  } else
  if constexpr (D == 2) {
    if (c->plastic) { if (e->W == 1) launch(k_assemble_nh<2, 1, true>, 1); else launch(k_assemble_nh<2, 2, true>, 2); }
    else { if (e->W == 1) launch(k_assemble_nh<2, 1, false>, 1); else launch(k_assemble_nh<2, 2, false>, 2); }
  } else {
    if (c->plastic) { if (e->W == 4) launch(k_assemble_nh<3, 4, true>, 4); else launch(k_assemble_nh<3, 8, true>, 8); }
    else { if (e->W == 4) launch(k_assemble_nh<3, 4, false>, 4); else launch(k_assemble_nh<3, 8, false>, 8); }
  }
  k_bsr_diag<D><<<nblk(e->max_act, 128), 128, 0, e->stream>>>(e->G, c->row_ptr, c->cols, c->vals, c->fx, c->a1, c->diag, c->own);
  if (e->slab_on) {
    imp_band_sum<D>(e, c->diag);
    k_diag_guard<<<IMP_NPART, 256, 0, e->stream>>>(e->G.n_active, D, c->diag);
  }
  e->launches += 2;
  c->assemblies++;
  return 0;
}

// The same iteration over several slabs (SURVEY 8e "Implicit").  Per iteration: the local product with this slab's share
// of the tangent, the band sums of the result (the force exchange of the explicit scheme: peer-memory stores or
// ncclSend/ncclRecv), one all-reduce of the p.Ap partials and one of the (r.z | r.r) partials -- no host round trip
// inside an iteration; every slab then re-sums the same reduced partial arrays in the same order, so alpha, beta and
// the convergence decision are identical everywhere.
template <int D>
static int imp_pcg_slabs_t(nlps_engine* e, const double* b, double* x) {
  ImplicitCtx* c = e->imp;
  const int* na = e->G.n_active;
  double* part_pAp = c->part + 4 * IMP_NPART;
  double* pair[2] = {c->pair, c->pair + 2 * IMP_NPART};  // [r.z | r.r] by parity
  k_pcg_init<<<IMP_NPART, 256, 0, e->stream>>>(na, D, b, c->diag, x, c->r, c->z, c->p, pair[0], pair[0] + IMP_NPART, c->own);
  if (imp_allreduce(e, pair[0], 2 * IMP_NPART)) return -1;
  auto norm_rr = [&](int q) {
    cudaMemcpyAsync(c->h_part, pair[q] + IMP_NPART, sizeof(double) * IMP_NPART, cudaMemcpyDeviceToHost, e->stream);
    cudaStreamSynchronize(e->stream);
    double s_ = 0.0;
    for (int i = 0; i < IMP_NPART; i++) s_ += c->h_part[i];
    return sqrt(s_);
  };
  const double bnorm = norm_rr(0);
  if (bnorm == 0.0) return 0;
  const double target = c->prm.pcg_rtol * bnorm;
  int it = 0;
  const int check = 8;
  while (it < c->prm.pcg_max_iter) {
    for (int k = 0; k < check; k++, it++) {
      const int q = it & 1;
      k_bsr_spmv<D><<<IMP_NSPMV, 256, 0, e->stream>>>(e->G, c->row_ptr, c->cols, c->vals, c->fx, c->a1, c->p, c->Ap, c->p, part_pAp, nullptr,
                                                      c->own);
      if (imp_band_sum<D>(e, c->Ap) || imp_allreduce(e, part_pAp, IMP_NSPMV)) return -it - 1;
      k_pcg_update1<<<IMP_NPART, 256, 0, e->stream>>>(na, D, part_pAp, pair[q], c->p, c->Ap, c->diag, x, c->r, c->z, pair[q ^ 1],
                                                     pair[q ^ 1] + IMP_NPART, c->own);
      if (imp_allreduce(e, pair[q ^ 1], 2 * IMP_NPART)) return -it - 1;
      k_pcg_update2<<<IMP_NPART, 256, 0, e->stream>>>(na, D, pair[q], pair[q ^ 1], c->z, c->p);
    }
    e->launches += 3 * check;
    const double rn = norm_rr(it & 1);
    if (!(rn == rn)) return -it;  // NaN: breakdown
    if (rn <= target) { c->pcg_iters += it; return it; }
  }
  c->pcg_iters += it;
  return -it;
}

// Jacobi-PCG on (K + alpha_1 M) delta = b; returns the iteration count (negative: not converged)
template <int D>
static int imp_pcg_t(nlps_engine* e, const double* b, double* x) {
  ImplicitCtx* c = e->imp;
  const int* na = e->G.n_active;
  double* part_pAp = c->part + 4 * IMP_NPART;
  double* part_rz[2] = {c->part, c->part + IMP_NPART};
  double* part_rr = c->part + 2 * IMP_NPART;
  if (e->slab_on) return imp_pcg_slabs_t<D>(e, b, x);
  k_pcg_init<<<IMP_NPART, 256, 0, e->stream>>>(na, D, b, c->diag, x, c->r, c->z, c->p, part_rz[0], part_rr);
  const double bnorm = imp_norm(e, 2);
  if (bnorm == 0.0) return 0;
  const double target = c->prm.pcg_rtol * bnorm;
  int it = 0;
  const int check = 8;
  while (it < c->prm.pcg_max_iter) {
    for (int k = 0; k < check; k++, it++) {
      k_bsr_spmv<D><<<IMP_NSPMV, 256, 0, e->stream>>>(e->G, c->row_ptr, c->cols, c->vals, c->fx, c->a1, c->p, c->Ap, c->p, part_pAp, nullptr);
      k_pcg_update1<<<IMP_NPART, 256, 0, e->stream>>>(na, D, part_pAp, part_rz[it & 1], c->p, c->Ap, c->diag, x, c->r, c->z,
                                                     part_rz[(it + 1) & 1], part_rr);
      k_pcg_update2<<<IMP_NPART, 256, 0, e->stream>>>(na, D, part_rz[it & 1], part_rz[(it + 1) & 1], c->z, c->p);
    }
    e->launches += 3 * check;
    const double rn = imp_norm(e, 2);
    if (!(rn == rn)) return -it;  // NaN: breakdown
    if (rn <= target) { c->pcg_iters += it; return it; }
  }
  c->pcg_iters += it;
  return -it;
}

template <int D>
static int imp_bicgstab_t(nlps_engine* e, const double* b, double* x) {
  ImplicitCtx* c = e->imp;
  const int* na = e->G.n_active;
  BiParts P_;
  P_.rho[0] = c->part2; P_.rho[1] = c->part2 + IMP_NPART; P_.rr = c->part2 + 2 * IMP_NPART;
  double* big = c->part2 + 3 * IMP_NPART;
  for (int q = 0; q < 2; q++) { P_.rv[q] = big + (size_t)q * IMP_NSPMV; P_.ts[q] = big + (size_t)(2 + q) * IMP_NSPMV; P_.tt[q] = big + (size_t)(4 + q) * IMP_NSPMV; }
  k_bi_init<<<IMP_NPART, 256, 0, e->stream>>>(na, D, b, x, c->r, c->brh, P_.rho[0], P_.rr);
  auto norm_rr = [&]() {
    cudaMemcpyAsync(c->h_part, P_.rr, sizeof(double) * IMP_NPART, cudaMemcpyDeviceToHost, e->stream);
    cudaStreamSynchronize(e->stream);
    double s_ = 0.0;
    for (int i = 0; i < IMP_NPART; i++) s_ += c->h_part[i];
    return sqrt(s_);
  };
  const double bnorm = norm_rr();
  if (bnorm == 0.0) return 0;
  const double target = c->prm.pcg_rtol * bnorm;
  int it = 0;
  const int check = 4;
  while (it < c->prm.pcg_max_iter) {
    for (int k = 0; k < check; k++, it++) {
      const int q = it & 1;
      k_bi_p<<<IMP_NPART, 256, 0, e->stream>>>(na, D, P_, it, c->r, c->bv, c->diag, c->p, c->by);
      k_bsr_spmv<D><<<IMP_NSPMV, 256, 0, e->stream>>>(e->G, c->row_ptr, c->cols, c->vals, c->fx, c->a1, c->by, c->bv, c->brh, P_.rv[q], nullptr);
      k_bi_s<<<IMP_NPART, 256, 0, e->stream>>>(na, D, P_, it, c->r, c->bv, c->diag, c->bs, c->z);
      k_bsr_spmv<D><<<IMP_NSPMV, 256, 0, e->stream>>>(e->G, c->row_ptr, c->cols, c->vals, c->fx, c->a1, c->z, c->bt, c->bs, P_.ts[q], P_.tt[q]);
      k_bi_x<<<IMP_NPART, 256, 0, e->stream>>>(na, D, P_, it, c->by, c->z, c->bs, c->bt, c->brh, x, c->r);
    }
    e->launches += 5 * check;
    const double rn = norm_rr();
    if (!(rn == rn)) return -it;
    if (rn <= target) { c->pcg_iters += it; return it; }
  }
  c->pcg_iters += it;
  return -it;
}

template <int D>
static int imp_step_t(nlps_engine* e, int step) {
  ImplicitCtx* c = e->imp;
  const int* na = e->G.n_active;
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0); cudaEventCreate(&t1);
  auto tick = [&]() { cudaEventRecord(t0, e->stream); };
  auto tock = [&](double& acc) { cudaEventRecord(t1, e->stream); cudaEventSynchronize(t1); float ms = 0; cudaEventElapsedTime(&ms, t0, t1); acc += ms; };
  imp_begin_t<D>(e, step);
  tick();
  imp_residual_t<D>(e, step, c->dU, c->R, 3);
  double rn = imp_norm(e, 3);
  tock(c->ms_residual);
  if (imp_poll(e)) return 1;
  c->res0 = rn;
  c->newton_iters = 0;
  const double tol = c->prm.tol;
  int status = 0;
  while (c->newton_iters < c->prm.max_iter && !(rn <= 100 * tol) && !(rn <= tol * c->res0)) {
    tick();
    imp_assemble_t<D>(e);
    tock(c->ms_assemble);
    k_vec_neg<<<IMP_NPART, 256, 0, e->stream>>>(na, D, c->R, c->Rt);  // Rt doubles as the right-hand side
    tick();
    const int its = c->plastic ? imp_bicgstab_t<D>(e, c->Rt, c->delta) : imp_pcg_t<D>(e, c->Rt, c->delta);
    tock(c->ms_pcg);
    if (imp_poll(e)) { status = 1; break; }
    if (its < 0 && getenv("NLPS_VERBOSE")) fprintf(stderr, "nlps_b200 (implicit): PCG stopped after %d iterations\n", -its);
    // step halving on |R| (the reference: SNES backtracking line search)
    tick();
    double lam = 1.0, rt = 0.0;
    bool ok = false;
    for (int ls = 0; ls < 8; ls++, lam *= 0.5) {
      k_vec_axpy<<<IMP_NPART, 256, 0, e->stream>>>(na, D, c->dU, lam, c->delta, c->trial);
      imp_residual_t<D>(e, step, c->trial, c->Rt, 3);
      rt = imp_norm(e, 3);
      if (rt < rn) { ok = true; break; }
    }
    if (!ok) {
      k_vec_axpy<<<IMP_NPART, 256, 0, e->stream>>>(na, D, c->dU, 1.0, c->delta, c->trial);
      imp_residual_t<D>(e, step, c->trial, c->Rt, 3);
      rt = imp_norm(e, 3);
    }
    tock(c->ms_residual);
    if (imp_poll(e)) { status = 1; break; }
    std::swap(c->dU, c->trial);
    std::swap(c->R, c->Rt);
    c->newton_iters++;
    const bool stalled = !ok && !(rt < rn);
    rn = rt;
    if (stalled) break;
  }
  c->res = rn;
  cudaEventDestroy(t0); cudaEventDestroy(t1);
  if (status) return status;
  // G3 + K4
  auto g2p = [&](auto kfn) {
    kfn<<<nblk(std::max(e->np, 1), 128), 128, 0, e->stream>>>(e->mesh, e->P, e->G, c->a1, c->a2, c->a3, c->a4, c->a5, c->a6, c->dU, c->Vn, c->An, e->alme);
  };
  if constexpr (D == 2) { if (e->W == 1) g2p(k_g2p_implicit<2, 1>); else g2p(k_g2p_implicit<2, 2>); }
  else { if (e->W == 4) g2p(k_g2p_implicit<3, 4>); else g2p(k_g2p_implicit<3, 8>); }
  e->launches++;
  if (!e->inert_synced) {
    if (e->np) k_sync_inert<D><<<nblk(e->np, 256), 256, 0, e->stream>>>(e->P, e->mat);
    e->inert_synced = 1;
  }
  e->n1_stale = 1;  // downloads hand the rolled state to the *_n1 host fields, as the reference's copy roll leaves them
  std::swap(e->P.F_n, e->P.F_n1);
  std::swap(e->P.J_n, e->P.J_n1);
  std::swap(e->P.be_n, e->P.be_n1);
  std::swap(e->P.eps_n, e->P.eps_n1);
  std::swap(e->P.kap_n, e->P.kap_n1);
  return imp_poll(e);
}

extern "C" {

int nlps_b200_newmark_setup(nlps_engine* e, const nlps_newmark* prm) {
  cudaSetDevice(e->device);
  if (!prm || (!prm->quasi_static && !(prm->beta > 0.0))) return 1;
  for (int i = 0; i < MAX_MATERIALS; i++)
    if (e->mat.m[i].type == NLPS_MAT_LADE_DUNCAN) {  // (no reference trace of a Lade-Duncan cloud exists to pin it to)
      fprintf(stderr, "nlps_b200_newmark_setup: the implicit scheme has tangents for Neo-Hookean, Hencky, Drucker-Prager, Matsuoka-Nakai and Von-Mises only\n");
      return 1;
    }
  return imp_setup(e, prm);
}

int nlps_b200_newmark_step(nlps_engine* e, int time_step) {
  cudaSetDevice(e->device);
  if (!e->imp) return 1;
  return e->D == 2 ? imp_step_t<2>(e, time_step) : imp_step_t<3>(e, time_step);
}

int nlps_b200_newmark_run(nlps_engine* e, int first_step, int count) {
  for (int k = first_step; k < first_step + count; k++)
    if (nlps_b200_newmark_step(e, k)) return 1;
  return 0;
}

int nlps_b200_newmark_stats(nlps_engine* e, nlps_newmark_stats* out) {
  if (!e->imp || !out) return 1;
  ImplicitCtx* c = e->imp;
  out->newton_iters = c->newton_iters;
  out->pcg_iters_total = c->pcg_iters;
  out->assemblies_total = c->assemblies;
  out->residual_evals_total = c->residual_evals;
  out->residual0 = c->res0;
  out->residual = c->res;
  out->ms_assemble = c->ms_assemble;
  out->ms_pcg = c->ms_pcg;
  out->ms_residual = c->ms_residual;
  int h[4] = {0, 0, 0, 0};
  cudaMemcpy(h, c->tops, sizeof(int) * 3, cudaMemcpyDeviceToHost);
  out->nnz_blocks = h[2];
  cudaMemcpy(h, e->G.n_active, sizeof(int), cudaMemcpyDeviceToHost);
  out->n_rows = h[0];
  return 0;
}

// ---- stage-level entry points (parity tests)
int nlps_b200_newmark_begin(nlps_engine* e, int time_step) {
  cudaSetDevice(e->device);
  if (!e->imp) return 1;
  if (e->D == 2) imp_begin_t<2>(e, time_step); else imp_begin_t<3>(e, time_step);
  return poll_error(e);
}

/* which: 0 v_n, 1 a_n, 2 dU (current iterate), 3 residual of the last evaluation; out is n_nodes x ndim */
int nlps_b200_newmark_get(nlps_engine* e, int which, double* out) {
  cudaSetDevice(e->device);
  if (!e->imp || which < 0 || which > 3) return 1;
  ImplicitCtx* c = e->imp;
  const double* src[4] = {c->Vn, c->An, c->dU, c->R};
  const size_t n = (size_t)e->nn * e->D;
  double* tmp = nullptr;
  CUDA_OK(cudaMalloc(&tmp, n * sizeof(double)));
  k_imp_export<<<nblk(n, 256), 256, 0, e->stream>>>(e->G, e->nn, e->D, src[which], tmp);
  cudaError_t st = cudaMemcpyAsync(out, tmp, n * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
  if (st == cudaSuccess) st = cudaStreamSynchronize(e->stream);
  cudaFree(tmp);
  return st == cudaSuccess ? 0 : 1;
}

/* residual at the nodal increment dU (n_nodes x ndim, full-grid indexing; NULL = the current iterate) */
int nlps_b200_newmark_residual(nlps_engine* e, int time_step, const double* dU, double* R) {
  cudaSetDevice(e->device);
  if (!e->imp) return 1;
  ImplicitCtx* c = e->imp;
  const size_t n = (size_t)e->nn * e->D;
  if (dU) {
    double* tmp = nullptr;
    CUDA_OK(cudaMalloc(&tmp, n * sizeof(double)));
    CUDA_OK(cudaMemcpyAsync(tmp, dU, n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    k_imp_import<<<nblk(n, 256), 256, 0, e->stream>>>(e->G, e->nn, e->D, tmp, c->dU);
    CUDA_OK(cudaStreamSynchronize(e->stream));
    cudaFree(tmp);
  }
  if (e->D == 2) imp_residual_t<2>(e, time_step, c->dU, c->R, 3); else imp_residual_t<3>(e, time_step, c->dU, c->R, 3);
  if (poll_error(e)) return 1;
  return R ? nlps_b200_newmark_get(e, 3, R) : 0;
}

/* Tangent of the state left by the last residual evaluation, block CSR over the active nodes WITHOUT the
 * alpha_1 M term and the Dirichlet treatment (both live in the operator).  Two calls: sizes (arrays NULL), fill.
 * row_nodes[n_rows], row_ptr[n_rows+1], col_nodes[nnz_blocks] are node ids, vals[nnz_blocks*d*d] row-major blocks. */
int nlps_b200_newmark_tangent(nlps_engine* e, int* n_rows, int* nnz_blocks, int* row_nodes, int* row_ptr, int* col_nodes,
                              double* vals) {
  cudaSetDevice(e->device);
  if (!e->imp) return 1;
  ImplicitCtx* c = e->imp;
  int h[4];
  CUDA_OK(cudaMemcpy(h, c->tops, sizeof(int) * 3, cudaMemcpyDeviceToHost));
  const int nnz = h[2];
  CUDA_OK(cudaMemcpy(h, e->G.n_active, sizeof(int), cudaMemcpyDeviceToHost));
  const int nr = h[0];
  if (n_rows) *n_rows = nr;
  if (nnz_blocks) *nnz_blocks = nnz;
  if (!row_ptr || !col_nodes || !vals || !row_nodes) return 0;
  if (e->D == 2) imp_assemble_t<2>(e); else imp_assemble_t<3>(e);
  if (poll_error(e)) return 1;
  int* tmp = nullptr;
  CUDA_OK(cudaMalloc(&tmp, sizeof(int) * std::max(nnz, 1)));
  k_csr_cols_to_nodes<<<nblk(std::max(nnz, 1), 256), 256, 0, e->stream>>>(e->G, c->cols, nnz, tmp);
  CUDA_OK(cudaMemcpyAsync(col_nodes, tmp, sizeof(int) * nnz, cudaMemcpyDeviceToHost, e->stream));
  CUDA_OK(cudaMemcpyAsync(row_ptr, c->row_ptr, sizeof(int) * (nr + 1), cudaMemcpyDeviceToHost, e->stream));
  CUDA_OK(cudaMemcpyAsync(row_nodes, e->G.act_list, sizeof(int) * nr, cudaMemcpyDeviceToHost, e->stream));
  CUDA_OK(cudaMemcpyAsync(vals, c->vals, sizeof(double) * (size_t)nnz * e->D * e->D, cudaMemcpyDeviceToHost, e->stream));
  CUDA_OK(cudaStreamSynchronize(e->stream));
  cudaFree(tmp);
  return 0;
}

/* The whole scheme call with HOST buffers (what U_Newmark_Beta does for the driver, U-Newmark-beta.c:130-425). */
int nlps_b200_u_newmark_beta(const nlps_mesh* mesh, const nlps_solver* solver, const nlps_newmark* newmark, int n_bounds,
                             const nlps_load* bounds, int n_neumann, const nlps_load* neumann, const double* gravity,
                             int n_materials, const nlps_material* materials, nlps_particles* state, int run_initialize,
                             int results_every, nlps_results_cb cb, void* user, int device) {
  char msg[256];
  nlps_engine* e = nlps_b200_create(mesh, solver, n_bounds, bounds, n_neumann, neumann, gravity, n_materials, materials, state,
                                    device, msg, sizeof(msg));
  if (!e) return 1;
  int status = nlps_b200_newmark_setup(e, newmark);
  if (!status && run_initialize) status = nlps_b200_initialize_lme(e);
  for (int k = solver->initial_step; !status && k < solver->num_steps; k++) {
    status = nlps_b200_newmark_step(e, k);
    if (!status && results_every > 0 && k % results_every == 0) {  // U-Newmark-beta.c:409-411
      status = nlps_b200_download(e, state);
      if (!status && cb) cb(k, user);
    }
  }
  if (!status) status = nlps_b200_download(e, state);
  nlps_b200_destroy(e);
  return status;
}

}  // extern "C"
