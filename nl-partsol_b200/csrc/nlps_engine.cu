// nlps_engine.cu -- B200-native (sm_100a) engine for NL-PartSol's explicit NPC-FS step.
//
// Design: fp64 everywhere, particle state SoA in HBM, mesh adjacency as
// CSR in the reference's chain order.  Neighbour lists are stored as BITMASKS over the
// 2-ring of the closest node (4 B / 16 B per particle instead of 4n B).  Particle-to-grid
// assembly is an ATOMICS-FREE, cell-sorted gather: particles are binned by closest node
// (I0) every step, and one thread per active node sums the contributions of the particles
// of the cells in its 2-ring (deterministic order, no fp64 atomics -- shared-memory fp64
// atomicAdd is a CAS loop on sm_100a).  Grid update + Dirichlet BCs are fused into the
// node kernels; kinematics + stress update + the per-particle force operator are one
// particle kernel; state roll is a pointer swap.
//
// Reference citations are relative to nl-partsol/src of migmolper/NL-PartSol.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/nlps_b200.h"
#include "nlps_device.cuh"

#define CUDA_OK(call)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      fprintf(stderr, "nlps_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

static const int MAX_MATERIALS = 8;
static const int MAX_MASK_WORDS = 8;  // 2-ring up to 256 nodes
__constant__ MatParams c_mat[MAX_MATERIALS];

enum KernelId {
  K_SEARCH = 0, K_NODE_FLAGS, K_SCAN1, K_SCAN2, K_SCAN3, K_FILL, K_NODE_FINISH, K_LME, K_P2G_MASS_DISP,
  K_GRID_DISP, K_KIN_STRESS, K_TRACTION, K_P2G_FORCE, K_GRID_ACC, K_G2P, K_COUNT
};
static const char* kKernelNames[K_COUNT] = {
    "search_closest_node", "node_flags", "scan_reduce", "scan_tops", "scan_apply", "cell_fill",
    "node_finish", "lme_update", "p2g_mass_disp", "grid_disp_bc", "g2p_kin_stress", "traction", "p2g_force",
    "grid_acc", "g2p_update"};

// ---------------------------------------------------------------------------
// Device views
// node records are padded so that one node is one or two 16-byte vector loads
template <int D> struct NS { static constexpr int X = (D == 2) ? 2 : 4; };   // coordinates stride (doubles)
template <int D>
__device__ __forceinline__ void ldvec(const double* p, double* out) {
  double2 a = *reinterpret_cast<const double2*>(p);
  out[0] = a.x; out[1] = a.y;
  if (D == 3) { double2 b = *reinterpret_cast<const double2*>(p + 2); out[2] = b.x; }
}
struct MeshDev {
  int nn;
  const double* X;  // nn x NS<D>::X (row-major, padded)
  const int *r1p, *r1i, *r2p, *r2i;
  const int *r1tp, *r1ti, *r2tp, *r2ti;  // transposed adjacency (who lists me)
  const unsigned char* r2q;              // r2q[r2p[B]+s] = position of B inside the r2t row of node r2i[r2p[B]+s]
  const double* h_avg;
};

// AoS record read by the node-centric gather kernels.  Layout (doubles):
// [0..D) x, [D] sstar, [D+1] beta, [D+2..2D+2) lambda, [2D+2] zinv, [2D+3] mass,
// [2D+4..3D+4) D_dis, [3D+4 .. 3D+4+D*D) G (force operator), then D traction*area.
template <int D>
struct Rec {
  static constexpr int X = 0, SSTAR = D, BETA = D + 1, LAM = D + 2, ZINV = 2 * D + 2, MASS = 2 * D + 3,
                       DDIS = 2 * D + 4, G = 3 * D + 4, TRAC = 3 * D + 4 + D * D,
                       SIZE = ((3 * D + 4 + D * D + D) + 1) & ~1;
};

struct PartDev {
  int np;
  // SoA, component-major: f[c*np + p]
  double *x, *dis, *ddis, *vel, *acc, *lam;
  double *beta, *mass, *vol0, *rho, *W;
  double *J_n, *J_n1, *eps_n, *eps_n1, *kap_n, *kap_n1;
  double *F_n, *F_n1, *DF, *be_n, *be_n1, *stress, *cep;
  double *Fs4, *DFs4;  // 2D slot 4 of F / DF (never touched by the kinematics, Appendix B)
  double* rec;
  int *I0, *nnodes, *matidx;
  uint32_t* mask;  // W words, word-major: mask[w*np + p]
};

struct GridDev {
  double *M, *F;  // M: nn ; F: nn x D (row-major)
  double* UA;     // per node [dU (NS) | A (NS)]: the two nodal fields the G2P gathers read, one record
  unsigned char *active, *fixed;
  int *cnt, *cursor, *cell_start, *plist, *act_list, *n_active;
  int *occ_list, *n_occ, *act_pos, *occ_pos;
  ulonglong2 *packed, *scan_blk;
  double* part;  // per (active node, r2t slot): (1+D) partial sums written by the cell kernels
  int cap;
};

struct StepParams {
  double dt, gamma_lme, neg_log_tol, tol_wrapper, thickness;
  int max_iter_lme, nsteps, step, update_I0, W;
  ReturnMapParams rp;
};

// ---------------------------------------------------------------------------
__device__ __forceinline__ void latch_error(int* err, int code, int p) {
  if (atomicCAS(&err[0], 0, code) == 0) err[1] = p;
}

// squared distance with the reference's rounding sequence: sum_i (x_i - X_i)*(x_i - X_i),
// products and sums rounded separately (no FMA contraction), Nodes-Tools.c:400-420 and
// MatrixOp.c:895-920.  Needed for bit-exact closest node / neighbour lists.
template <int D>
__device__ __forceinline__ double dist2_exact(const double* xp, const double* XA, double* l) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < D; i++) {
    l[i] = __dsub_rn(xp[i], XA[i]);
    s = __dadd_rn(s, __dmul_rn(l[i], l[i]));
  }
  return s;
}

// largest s with sqrt_rn(s) <= Ra, so that "s <= sstar" is EXACTLY the reference's
// "sqrt(s) <= Ra" (LME.c:1074) without a square root per candidate.
__device__ inline double sstar_from_Ra(double Ra) {
  if (!(Ra < 1.0e150)) return (Ra != Ra) ? -1.0 : 1.0e300;
  double t = __dmul_rn(Ra, Ra);
  for (int it = 0; it < 4 && __dsqrt_rn(t) > Ra; it++) t = __longlong_as_double(__double_as_longlong(t) - 1);
  for (int it = 0; it < 4; it++) {
    double u = __longlong_as_double(__double_as_longlong(t) + 1);
    if (__dsqrt_rn(u) <= Ra) t = u; else break;
  }
  return t;
}


// Neighbour iteration with 4-way memory-level parallelism: the kernels are bound by the latency of the
// dependent gathers (mask bit -> node id -> node data), so node ids and node data of GROUPS of four
// neighbours are requested before any of them is consumed.
#define NLPS_GROUP 4
template <int W>
__device__ __forceinline__ int next_group(const MeshDev& m, int base, uint32_t* mk, int& w, int* node) {
  int cnt = 0;
#pragma unroll
  for (int u = 0; u < NLPS_GROUP; u++) {
    while (w < W && mk[w] == 0u) w++;
    if (w < W) {
      int b = __ffs(mk[w]) - 1;
      mk[w] &= mk[w] - 1;
      node[u] = m.r2i[base + w * 32 + b];
      cnt = u + 1;
    } else {
      node[u] = -1;
    }
  }
  return cnt;
}

// ---------------------------------------------------------------------------
// K0a: closest node + cell histogram.   local_search__LME__ first loop (LME.c:917-944),
// get_closest_node__MeshTools__ (Nodes-Tools.c:476-538): first strict minimum, chain order.
template <int D>
__global__ void __launch_bounds__(256) k_search(MeshDev m, PartDev P, GridDev G, int update_I0) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.np) return;
  int I0 = P.I0[p];
  if (update_I0) {
    double nd = 0.0, xp[D];
#pragma unroll
    for (int i = 0; i < D; i++) {
      double dd = P.dis[i * P.np + p];
      nd = __dadd_rn(nd, __dmul_rn(dd, dd));
      xp[i] = P.x[i * P.np + p];
    }
    if (nd > 0.0) {  // norm__MatrixLib__(dis) > 0  (LME.c:924)
      int b0 = m.r1p[I0], b1 = m.r1p[I0 + 1];
      double dmin = 0.0, l[D];
      int best = I0;
      for (int q = b0; q < b1; q++) {
        int node = m.r1i[q];
        double dq = __dsqrt_rn(dist2_exact<D>(xp, &m.X[(size_t)node * NS<D>::X], l));
        if (q == b0 || dq < dmin) { dmin = dq; best = node; }
      }
      I0 = best;
      P.I0[p] = I0;
    }
  }
  atomicAdd(&G.cnt[I0], 1);
}

// node flags: ActiveNode[A] = OR over particles with I0 in {B : A in ring1(B)} (LME.c:949-965,
// after the reset of Shape-Functions.c:38-47).  packed.x = cnt | occupied << 40, packed.y = active:
// one fused exclusive scan yields cell_start, the rank of each occupied cell and of each active node.
__global__ void __launch_bounds__(256) k_node_flags(MeshDev m, GridDev G) {
  int A = blockIdx.x * blockDim.x + threadIdx.x;
  if (A >= m.nn) return;
  int act = 0;
  for (int q = m.r1tp[A]; q < m.r1tp[A + 1] && !act; q++) act = G.cnt[m.r1ti[q]] > 0;
  G.active[A] = (unsigned char)act;
  unsigned long long c = (unsigned)G.cnt[A];
  G.packed[A] = make_ulonglong2(c | ((unsigned long long)(c > 0) << 40), (unsigned long long)act);
  G.cursor[A] = 0;
}

__device__ __forceinline__ ulonglong2 add2(ulonglong2 a, ulonglong2 b) { return make_ulonglong2(a.x + b.x, a.y + b.y); }

// exclusive scan of packed, 3 phases, 2048 items per block
static const int SCAN_ITEMS = 2048;
__global__ void __launch_bounds__(256) k_scan_reduce(const ulonglong2* in, ulonglong2* blk, int n) {
  __shared__ ulonglong2 sh[256];
  size_t base = (size_t)blockIdx.x * SCAN_ITEMS;
  ulonglong2 s = make_ulonglong2(0, 0);
  for (int i = threadIdx.x; i < SCAN_ITEMS; i += 256)
    if (base + i < (size_t)n) s = add2(s, in[base + i]);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] = add2(sh[threadIdx.x], sh[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) blk[blockIdx.x] = sh[0];
}
__global__ void k_scan_tops(ulonglong2* blk, int nblk, int* n_active, int* n_occ, int* n_particles_check) {
  // single thread block, serial over chunks (nblk is a few thousand)
  __shared__ ulonglong2 sh[1024];
  __shared__ ulonglong2 carry;
  if (threadIdx.x == 0) carry = make_ulonglong2(0, 0);
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    int i = base + threadIdx.x;
    ulonglong2 v = (i < nblk) ? blk[i] : make_ulonglong2(0, 0);
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      ulonglong2 t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : make_ulonglong2(0, 0);
      __syncthreads();
      sh[threadIdx.x] = add2(sh[threadIdx.x], t);
      __syncthreads();
    }
    if (i < nblk) {
      ulonglong2 c = carry, inc = sh[threadIdx.x];
      blk[i] = make_ulonglong2(c.x + inc.x - v.x, c.y + inc.y - v.y);
    }
    __syncthreads();
    if (threadIdx.x == 0) carry = add2(carry, sh[1023]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *n_active = (int)carry.y;
    *n_occ = (int)(carry.x >> 40);
    *n_particles_check = (int)(carry.x & 0xffffffffffull);
  }
}
__global__ void __launch_bounds__(256) k_scan_apply(const ulonglong2* in, const ulonglong2* blk, int* cell_start,
                                                    int* occ_pos, int* act_pos, int n) {
  __shared__ ulonglong2 sh[256];
  size_t base = (size_t)blockIdx.x * SCAN_ITEMS;
  const int per = SCAN_ITEMS / 256;
  ulonglong2 v[per], s = make_ulonglong2(0, 0);
#pragma unroll
  for (int k = 0; k < per; k++) {
    size_t i = base + (size_t)threadIdx.x * per + k;
    v[k] = (i < (size_t)n) ? in[i] : make_ulonglong2(0, 0);
    s = add2(s, v[k]);
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    ulonglong2 t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : make_ulonglong2(0, 0);
    __syncthreads();
    sh[threadIdx.x] = add2(sh[threadIdx.x], t);
    __syncthreads();
  }
  ulonglong2 b = blk[blockIdx.x], inc = sh[threadIdx.x];
  ulonglong2 run = make_ulonglong2(b.x + inc.x - s.x, b.y + inc.y - s.y);
#pragma unroll
  for (int k = 0; k < per; k++) {
    size_t i = base + (size_t)threadIdx.x * per + k;
    if (i < (size_t)n) {
      cell_start[i] = (int)(run.x & 0xffffffffffull);
      occ_pos[i] = (int)(run.x >> 40);
      act_pos[i] = (int)run.y;
    }
    run = add2(run, v[k]);
  }
}

__global__ void __launch_bounds__(256) k_cell_fill(PartDev P, GridDev G) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.np) return;
  int I0 = P.I0[p];
  int pos = G.cell_start[I0] + atomicAdd(&G.cursor[I0], 1);
  G.plist[pos] = p;
}

// per node: sort the cell's particle ids ascending (deterministic summation order) and
// append occupied cells / active nodes to their compact lists.
__global__ void __launch_bounds__(256) k_node_finish(MeshDev m, GridDev G) {
  int A = blockIdx.x * blockDim.x + threadIdx.x;
  if (A >= m.nn) return;
  int n = G.cnt[A];
  if (n > 1) {
    int* a = G.plist + G.cell_start[A];
    for (int i = 1; i < n; i++) {
      int v = a[i], j = i - 1;
      while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; j--; }
      a[j + 1] = v;
    }
  }
  if (n > 0) G.occ_list[G.occ_pos[A]] = A;
  if (G.active[A]) G.act_list[G.act_pos[A]] = A;
}

// ---------------------------------------------------------------------------
// K0: per-particle LME update.  tributary__LME__ (LME.c:1019-1099) with the PREVIOUS beta,
// beta__LME__ (LME.c:177-185), __lambda_Newton_Rapson (LME.c:272-353), plus the explicit
// predictor (__predictor_PARTICLES, U-Verlet.c:229-253) and the gather record.
template <int D, int W>
__global__ void __launch_bounds__(128) k_lme(MeshDev m, PartDev P, GridDev G, StepParams sp, int* err,
                                             int do_predictor) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.np) return;
  const int np = P.np;
  const int I0 = P.I0[p];
  const int base = m.r2p[I0], len = m.r2p[I0 + 1] - base;
  double xp[D], lam[D];
#pragma unroll
  for (int i = 0; i < D; i++) { xp[i] = P.x[i * np + p]; lam[i] = P.lam[i * np + p]; }
  const double beta_old = P.beta[p];
  const double Ra = __dsqrt_rn(__ddiv_rn(sp.neg_log_tol, beta_old));  // LME.c:1052
  const double sstar = sstar_from_Ra(Ra);
  uint32_t mk[W];
#pragma unroll
  for (int w = 0; w < W; w++) mk[w] = 0u;
  int n = 0;
  for (int k = 0; k < len; k++) {
    int node = m.r2i[base + k];
    if (!G.active[node]) continue;
    double l[D];
    double s = dist2_exact<D>(xp, &m.X[(size_t)node * NS<D>::X], l);
    if (s <= sstar) { mk[k >> 5] |= 1u << (k & 31); n++; }
  }
#pragma unroll
  for (int w = 0; w < W; w++) P.mask[(size_t)w * np + p] = mk[w];
  P.nnodes[p] = n;
  if (n < D + 1) { latch_error(err, NLPS_ERR_FEW_NEIGHBOURS, p); return; }
  const double h = m.h_avg[I0];
  const double beta = __ddiv_rn(sp.gamma_lme, __dmul_rn(h, h));
  P.beta[p] = beta;

  // Newton on lambda
  int NumIter = 0;
  double Z = 1.0;
  bool failed = false;
  while (NumIter <= sp.max_iter_lme) {
    double r[D], JJ[D * D];
    Z = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) r[i] = 0.0;
#pragma unroll
    for (int i = 0; i < D * D; i++) JJ[i] = 0.0;
    {
      uint32_t mw[W];
#pragma unroll
      for (int w = 0; w < W; w++) mw[w] = mk[w];
      int wcur = 0, node[NLPS_GROUP];
      while (next_group<W>(m, base, mw, wcur, node) > 0) {
        double Xn[NLPS_GROUP][D];
#pragma unroll
        for (int u = 0; u < NLPS_GROUP; u++)
#pragma unroll
          for (int i = 0; i < D; i++) Xn[u][i] = 0.0;
#pragma unroll
        for (int u = 0; u < NLPS_GROUP; u++)
          if (node[u] >= 0) ldvec<D>(&m.X[(size_t)node[u] * NS<D>::X], Xn[u]);
#pragma unroll
        for (int u = 0; u < NLPS_GROUP; u++) {
          if (node[u] < 0) continue;
          double l[D], ll = 0.0, lx = 0.0;
#pragma unroll
          for (int i = 0; i < D; i++) {
            l[i] = xp[i] - Xn[u][i];
            ll += l[i] * l[i];
            lx += l[i] * lam[i];
          }
          double e = exp(-beta * ll + lx);
          Z += e;
#pragma unroll
          for (int i = 0; i < D; i++) {
            r[i] += e * l[i];
#pragma unroll
            for (int j = i; j < D; j++) JJ[i * D + j] += e * l[i] * l[j];
          }
        }
      }
    }
    double Zi = 1.0 / Z, nr = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) { r[i] *= Zi; nr += r[i] * r[i]; }
    nr = sqrt(nr);
    if (nr > sp.tol_wrapper) {
#pragma unroll
      for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = i; j < D; j++) {
          JJ[i * D + j] = JJ[i * D + j] * Zi - r[i] * r[j];
          JJ[j * D + i] = JJ[i * D + j];
        }
      if (rcond_as_reference<D>(JJ) < 1E-8) { failed = true; latch_error(err, NLPS_ERR_SINGULAR_HESSIAN, p); break; }
      double Ji[D * D];
      inverse<D>(JJ, Ji);
#pragma unroll
      for (int i = 0; i < D; i++) {
        double dl = 0.0;
#pragma unroll
        for (int j = 0; j < D; j++) dl += Ji[i * D + j] * r[j];
        lam[i] -= dl;
      }
      NumIter++;
    } else {
      break;
    }
  }
  if (!failed && NumIter >= sp.max_iter_lme) latch_error(err, NLPS_ERR_NEWTON_LME, p);
#pragma unroll
  for (int i = 0; i < D; i++) P.lam[i * np + p] = lam[i];

  // predictor + gather record
  double* rec = P.rec + (size_t)p * Rec<D>::SIZE;
  const double mp = P.mass[p];
#pragma unroll
  for (int i = 0; i < D; i++) {
    double dd;
    if (do_predictor) {
      double v = P.vel[i * np + p], a = P.acc[i * np + p];
      dd = sp.dt * v + 0.5 * (sp.dt * sp.dt) * a;
      P.ddis[i * np + p] = dd;
      P.vel[i * np + p] = v + (1 - 0.5) * sp.dt * a;  // gamma = 0.5, U-Verlet.c:76,248
    } else {
      dd = P.ddis[i * np + p];
    }
    rec[Rec<D>::X + i] = xp[i];
    rec[Rec<D>::LAM + i] = lam[i];
    rec[Rec<D>::DDIS + i] = dd;
  }
  rec[Rec<D>::SSTAR] = sstar;
  rec[Rec<D>::BETA] = beta;
  rec[Rec<D>::ZINV] = 1.0 / Z;
  rec[Rec<D>::MASS] = mp;
}

// ---------------------------------------------------------------------------
// Stage 1 (cell kernel): one WARP per occupied cell B (all particles with I0 == B), one LANE per
// node A of the 2-ring of B (IT lanes-rounds when the ring has more than 32 nodes).  The cell's particle
// records are staged into shared memory by the whole warp (coalesced 8-byte loads, one round trip for
// CH particles), then every lane walks them (shared-memory broadcasts): no scatter, no atomic.  The
// lane's partial sums over the cell go to part[(rank(A), slot of B in A's transposed row)].
//   FORCE = false: M_A, sum m_p N_A DU_p  (U-Verlet.c:166-225, 301-367)
//   FORCE = true : f_A = sum_p N_A (G_p l_A + t_p) == -V0 tau (DF^-T gradN_A) + N_A T A0
//                  (U-Newmark-beta.c:1257-1374, U-Verlet.c:805-902)
template <int D, int IT, bool FORCE>
__global__ void __launch_bounds__(128) k_p2g_cell(MeshDev m, PartDev P, GridDev G, int has_traction) {
  // CPW cells per warp: all their metadata, node data and particle records are requested before any
  // is consumed (the kernel is latency-bound: ~4 particles of work per cell).
  constexpr int CPW = (IT == 1) ? 4 : 1, CH = 8, SZ = Rec<D>::SIZE, NV = FORCE ? D : 1 + D;
  __shared__ double sm_all[4][CPW * CH * SZ];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w0 = (blockIdx.x * 4 + wib) * CPW;
  const int nocc = *G.n_occ;
  if (w0 >= nocc) return;
  double* sm = sm_all[wib];
  // lanes 0..CPW-1 fetch the metadata of one cell each
  int mB = 0, mc0 = 0, mn = 0, mbase = 0, mlen = 0;
  if (lane < CPW && w0 + lane < nocc) {
    mB = G.occ_list[w0 + lane];
    mc0 = G.cell_start[mB];
    mn = G.cnt[mB];
    mbase = m.r2p[mB];
    mlen = m.r2p[mB + 1] - mbase;
  }
  int c0[CPW], n[CPW], base[CPW], len[CPW];
#pragma unroll
  for (int c = 0; c < CPW; c++) {
    c0[c] = __shfl_sync(0xffffffffu, mc0, c);
    n[c] = __shfl_sync(0xffffffffu, mn, c);
    base[c] = __shfl_sync(0xffffffffu, mbase, c);
    len[c] = __shfl_sync(0xffffffffu, mlen, c);
  }
  double XA[CPW][IT][D], acc[CPW][IT][NV];
  long long dst[CPW][IT];
#pragma unroll
  for (int c = 0; c < CPW; c++)
#pragma unroll
    for (int it = 0; it < IT; it++) {
      const int s = lane + 32 * it;
      dst[c][it] = -1;
#pragma unroll
      for (int v = 0; v < NV; v++) acc[c][it][v] = 0.0;
#pragma unroll
      for (int i = 0; i < D; i++) XA[c][it][i] = 0.0;
      if (s < len[c]) {
        const int A = m.r2i[base[c] + s];
        if (G.active[A]) {
          dst[c][it] = ((long long)G.act_pos[A] * G.cap + m.r2q[base[c] + s]) * (1 + D);
#pragma unroll
          for (int i = 0; i < D; i++) XA[c][it][i] = m.X[(size_t)A * NS<D>::X + i];
        }
      }
    }
  int nmax = 0;
#pragma unroll
  for (int c = 0; c < CPW; c++) nmax = max(nmax, n[c]);
  for (int j0 = 0; j0 < nmax; j0 += CH) {
    __syncwarp();
#pragma unroll
    for (int c = 0; c < CPW; c++) {
      const int nc = min(CH, n[c] - j0);
      if (nc <= 0) continue;
      const int pid = (lane < nc) ? G.plist[c0[c] + j0 + lane] : 0;
      for (int e0 = 0; e0 < nc * SZ; e0 += 32) {
        const int e = e0 + lane, j = min(e / SZ, nc - 1);
        const int pj = __shfl_sync(0xffffffffu, pid, j);
        if (e < nc * SZ) sm[c * CH * SZ + e] = P.rec[(size_t)pj * SZ + (e - j * SZ)];
      }
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < CPW; c++) {
      const int nc = min(CH, n[c] - j0);
#pragma unroll
      for (int it = 0; it < IT; it++) {
        if (dst[c][it] < 0) continue;
        for (int j = 0; j < nc; j++) {
          const double* rec = sm + (c * CH + j) * SZ;
          double l[D];
          double s2 = dist2_exact<D>(rec + Rec<D>::X, XA[c][it], l);
          if (s2 <= rec[Rec<D>::SSTAR]) {
            double lx = 0.0;
#pragma unroll
            for (int i = 0; i < D; i++) lx += l[i] * rec[Rec<D>::LAM + i];
            double N = exp(-rec[Rec<D>::BETA] * s2 + lx) * rec[Rec<D>::ZINV];
            if (!FORCE) {
              double mN = N * rec[Rec<D>::MASS];
              acc[c][it][0] += mN;
#pragma unroll
              for (int i = 0; i < D; i++) acc[c][it][1 + i] += mN * rec[Rec<D>::DDIS + i];
            } else {
#pragma unroll
              for (int i = 0; i < D; i++) {
                double gl = 0.0;
#pragma unroll
                for (int k = 0; k < D; k++) gl += rec[Rec<D>::G + i * D + k] * l[k];
                if (has_traction) gl += rec[Rec<D>::TRAC + i];
                acc[c][it][i] += N * gl;
              }
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CPW; c++)
#pragma unroll
    for (int it = 0; it < IT; it++)
      if (dst[c][it] >= 0) {
#pragma unroll
        for (int v = 0; v < NV; v++) G.part[dst[c][it] + v] = acc[c][it][v];
      }
}

// Stage 2 + G1 (node kernel): M_A = sum_p N_A m_p (U-Verlet.c:166-225); DU_A = sum_p m_p N_A DU_p / M_A
// (U-Verlet.c:301-367) as a contiguous, fixed-order sum of the cell partials; Dirichlet overwrite
// (U-Verlet.c:458-526) and restricted-DOF flags (Nodes-Tools.c:70-156).
struct BcDev {
  const int *node_ptr, *node_bnd;  // CSR: node -> boundary ids in boundary order
  const int* bnd_dim;
  const int* dir;      // [b][k][step] flattened with stride maxdim*nsteps
  const double* val;
  int maxdim, nsteps, nb;
};

template <int D>
__global__ void __launch_bounds__(128) k_grid_disp(MeshDev m, GridDev G, BcDev bc, int step) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *G.n_active) return;
  const int A = G.act_list[t];
  double mom[D], M = 0.0;
#pragma unroll
  for (int i = 0; i < D; i++) mom[i] = 0.0;
  const int q0 = m.r2tp[A], nq = m.r2tp[A + 1] - q0;
  const double* src = G.part + (size_t)t * G.cap * (1 + D);
  for (int q = 0; q < nq; q++) {
    if (G.cnt[m.r2ti[q0 + q]] == 0) continue;
    M += src[q * (1 + D)];
#pragma unroll
    for (int i = 0; i < D; i++) mom[i] += src[q * (1 + D) + 1 + i];
  }
  double dU[D];
#pragma unroll
  for (int i = 0; i < D; i++) dU[i] = mom[i] / M;
  unsigned fx = 0;
  for (int q = bc.node_ptr[A]; q < bc.node_ptr[A + 1]; q++) {
    int b = bc.node_bnd[q];
    for (int k = 0; k < bc.bnd_dim[b] && k < D; k++) {
      size_t o = ((size_t)b * bc.maxdim + k) * bc.nsteps + step;
      if (bc.dir[o] == 1) {
#pragma unroll
        for (int i = 0; i < D; i++) if (i == k) dU[i] = bc.val[o];
        fx |= 1u << k;
      }
    }
  }
  G.M[A] = M;
#pragma unroll
  for (int i = 0; i < D; i++) G.UA[(size_t)A * 2 * NS<D>::X + i] = dU[i];
  G.fixed[A] = (unsigned char)fx;
}

// ---------------------------------------------------------------------------
// K2 (+ the particle half of K3): kinematics, stress, force operator.
// DF = I + sum_A DU_A (x) gradN_A with gradN_a = -p_a J^-1 l_a  (compute-Strains.c:20-44,
// LME.c:836-891); F_n1 = DF F_n (compute-Strains.c:76-105); J > 0 (U-Verlet.c:608-613);
// rho /= det DF (U-Verlet.c:630-632); stress (Constitutive.c:18-258);
// G_p = V0 tau DF^-T J^-1 so that f_A = sum_p N_A G_p l_A  ==  -V0 tau (DF^-T gradN_A)
// (U-Newmark-beta.c:1257-1374 with Shape-Functions.c:405-448).
// MAT: compile-time material law when every particle uses the same one (keeps the Matsuoka-Nakai
// Newton out of the register budget of the other laws); -1 = mixed, dispatched per particle.
template <int D, int W, int MAT>
__global__ void __launch_bounds__(128, (MAT == 0 || MAT == 1) ? 4 : 2) k_kin_stress(MeshDev m, PartDev P, GridDev G, StepParams sp, int* err) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.np) return;
  const int np = P.np;
  constexpr int T = (D == 2) ? 5 : 9;
  const int I0 = P.I0[p];
  const int base = m.r2p[I0];
  double xp[D], lam[D];
#pragma unroll
  for (int i = 0; i < D; i++) { xp[i] = P.x[i * np + p]; lam[i] = P.lam[i * np + p]; }
  const double beta = P.beta[p];
  double Z = 0.0, r[D], JJ[D * D], Bm[D * D];
#pragma unroll
  for (int i = 0; i < D; i++) r[i] = 0.0;
#pragma unroll
  for (int i = 0; i < D * D; i++) { JJ[i] = 0.0; Bm[i] = 0.0; }
  {
    uint32_t mw[W];
#pragma unroll
    for (int w = 0; w < W; w++) mw[w] = P.mask[(size_t)w * np + p];
    int wcur = 0, node[NLPS_GROUP];
    while (next_group<W>(m, base, mw, wcur, node) > 0) {
      double Xn[NLPS_GROUP][D], dUn[NLPS_GROUP][D];
#pragma unroll
      for (int u = 0; u < NLPS_GROUP; u++)
#pragma unroll
        for (int i = 0; i < D; i++) { Xn[u][i] = 0.0; dUn[u][i] = 0.0; }
#pragma unroll
      for (int u = 0; u < NLPS_GROUP; u++)
        if (node[u] >= 0) {
          ldvec<D>(&m.X[(size_t)node[u] * NS<D>::X], Xn[u]);
          ldvec<D>(&G.UA[(size_t)node[u] * 2 * NS<D>::X], dUn[u]);
        }
#pragma unroll
      for (int u = 0; u < NLPS_GROUP; u++) {
        if (node[u] < 0) continue;
        double l[D], ll = 0.0, lx = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) {
          l[i] = xp[i] - Xn[u][i];
          ll += l[i] * l[i];
          lx += l[i] * lam[i];
        }
        double e = exp(-beta * ll + lx);
        Z += e;
#pragma unroll
        for (int i = 0; i < D; i++) {
          r[i] += e * l[i];
#pragma unroll
          for (int j = 0; j < D; j++) {
            if (j >= i) JJ[i * D + j] += e * l[i] * l[j];
            Bm[i * D + j] += e * dUn[u][i] * l[j];
          }
        }
      }
    }
  }
  const double Zi = 1.0 / Z;
#pragma unroll
  for (int i = 0; i < D; i++) r[i] *= Zi;
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = i; j < D; j++) {
      JJ[i * D + j] = JJ[i * D + j] * Zi - r[i] * r[j];
      JJ[j * D + i] = JJ[i * D + j];
    }
  double Ji[D * D];
  inverse<D>(JJ, Ji);
  double DF[D * D], Fn[D * D], Fn1[D * D];
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = 0; j < D; j++) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++) s += Bm[i * D + k] * Ji[k * D + j];
      DF[i * D + j] = ((i == j) ? 1.0 : 0.0) - s * Zi;
      Fn[i * D + j] = P.F_n[(size_t)(i * D + j) * np + p];
    }
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = 0; j < D; j++) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++) s += DF[i * D + k] * Fn[k * D + j];
      Fn1[i * D + j] = s;
      P.F_n1[(size_t)(i * D + j) * np + p] = s;
      P.DF[(size_t)(i * D + j) * np + p] = DF[i * D + j];
    }
  const double J1 = det<D>(Fn1);
  P.J_n1[p] = J1;
  if (J1 <= 0.0) { latch_error(err, NLPS_ERR_NEGATIVE_JACOBIAN, p); return; }
  const double dJ = det<D>(DF);
  P.rho[p] = P.rho[p] / dJ;

  // constitutive update
  const MatParams& mat = c_mat[P.matidx[p]];
  double tau[T], Wp = 0.0;
  const int mtype = (MAT >= 0) ? MAT : mat.type;
  if (mtype == NLPS_MAT_NEO_HOOKEAN_WRIGGERS) {
    stress_neo_hookean<D>(mat, Fn1, J1, tau, Wp);
  } else {
    constexpr int TB = (D == 2) ? 5 : 9;
    double be[TB], cep[D * D];
#pragma unroll
    for (int i = 0; i < TB; i++) be[i] = P.be_n[(size_t)i * np + p];
    double eps = P.eps_n[p], kap = P.kap_n[p];
    int st;
    if (MAT == NLPS_MAT_DRUCKER_PRAGER) st = stress_drucker_prager<D>(mat, sp.rp, DF, be, eps, kap, tau, Wp, cep);
    else if (MAT == NLPS_MAT_MATSUOKA_NAKAI) st = stress_matsuoka_nakai<D>(mat, sp.rp, DF, be, eps, kap, tau, Wp, cep);
    else st = (mtype == NLPS_MAT_DRUCKER_PRAGER) ? stress_drucker_prager<D>(mat, sp.rp, DF, be, eps, kap, tau, Wp, cep)
                                                 : stress_matsuoka_nakai<D>(mat, sp.rp, DF, be, eps, kap, tau, Wp, cep);
    if (st != 0) { latch_error(err, st, p); return; }
#pragma unroll
    for (int i = 0; i < TB; i++) P.be_n1[(size_t)i * np + p] = be[i];
    P.eps_n1[p] = eps;
    P.kap_n1[p] = kap;
    if (sp.rp.want_cep)
#pragma unroll
      for (int i = 0; i < D * D; i++) P.cep[(size_t)i * np + p] = cep[i];
  }
#pragma unroll
  for (int i = 0; i < T; i++) P.stress[(size_t)i * np + p] = tau[i];
  P.W[p] = Wp;

  // force operator G = V0 * tau * DF^-T * J^-1
  double DFi[D * D];
  double dd = inverse<D>(DF, DFi);
  if (dd == 0.0) { latch_error(err, NLPS_ERR_SINGULAR_DF, p); return; }
  const double V0 = P.vol0[p];
  double tA[D * D];
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = 0; j < D; j++) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++) s += tau[i * D + k] * DFi[j * D + k];  // tau * DF^-T
      tA[i * D + j] = s;
    }
  double* rec = P.rec + (size_t)p * Rec<D>::SIZE;
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = 0; j < D; j++) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++) s += tA[i * D + k] * Ji[k * D + j];
      rec[Rec<D>::G + i * D + j] = V0 * s;
    }
}

// Neumann tractions: per loaded particle t_p = sum_loads T(step) * A0_p, A0 = Vol_0 / thickness
// in 2D (U-Verlet.c:826-869).  Written into the gather record (zero for unloaded particles).
struct NeuDev {
  int n_entries;        // flattened (load, particle) pairs, load-major
  const int* part;
  const int* load;
  const int* load_dim;
  const int* dir;
  const double* val;
  int maxdim, nsteps;
};
template <int D>
__global__ void k_traction_clear(PartDev P) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.np) return;
#pragma unroll
  for (int i = 0; i < D; i++) P.rec[(size_t)p * Rec<D>::SIZE + Rec<D>::TRAC + i] = 0.0;
}
template <int D>
__global__ void k_traction(PartDev P, NeuDev nu, double thickness, int step) {
  // serial over entries of one particle is not needed: entries of different loads may hit the same
  // particle, so accumulate with fp64 global atomics (tiny set).
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nu.n_entries) return;
  int p = nu.part[e], b = nu.load[e];
  double A0 = P.vol0[p] / thickness;
  for (int k = 0; k < nu.load_dim[b] && k < D; k++) {
    size_t o = ((size_t)b * nu.maxdim + k) * nu.nsteps + step;
    if (nu.dir[o] == 1) atomicAdd(&P.rec[(size_t)p * Rec<D>::SIZE + Rec<D>::TRAC + k], nu.val[o] * A0);
  }
}

// K3 stage 2 + G2 (node kernel): f_A = sum of cell partials; a_A = g + f_A / M_A on free DOFs, 0 on
// restricted ones (U-Verlet.c:947-958; gravity as U-Newmark-beta.c:1539-1543).
template <int D>
__global__ void __launch_bounds__(128) k_grid_acc(MeshDev m, GridDev G, const double* grav, int nsteps, int step) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *G.n_active) return;
  const int A = G.act_list[t];
  double f[D];
#pragma unroll
  for (int i = 0; i < D; i++) f[i] = 0.0;
  const int q0 = m.r2tp[A], nq = m.r2tp[A + 1] - q0;
  const double* src = G.part + (size_t)t * G.cap * (1 + D);
  for (int q = 0; q < nq; q++) {
    if (G.cnt[m.r2ti[q0 + q]] == 0) continue;
#pragma unroll
    for (int i = 0; i < D; i++) f[i] += src[q * (1 + D) + i];
  }
  const double M = G.M[A];
  const unsigned fx = G.fixed[A];
#pragma unroll
  for (int i = 0; i < D; i++) {
    double g = grav ? grav[(size_t)i * nsteps + step] : 0.0;
    G.F[(size_t)A * D + i] = f[i];
    G.UA[(size_t)A * 2 * NS<D>::X + NS<D>::X + i] = ((fx >> i) & 1u) ? 0.0 : g + f[i] / M;
  }
}

// K4: G2P + corrector (U-Verlet.c:963-1084).  The n+1 -> n roll of F, J, b_e, kappa, EPS is a
// pointer swap on the host side of the engine.
template <int D, int W>
__global__ void __launch_bounds__(128) k_g2p(MeshDev m, PartDev P, GridDev G, StepParams sp) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.np) return;
  const int np = P.np;
  const int I0 = P.I0[p];
  const int base = m.r2p[I0];
  double xp[D], lam[D], a[D], du[D];
#pragma unroll
  for (int i = 0; i < D; i++) { xp[i] = P.x[i * np + p]; lam[i] = P.lam[i * np + p]; a[i] = 0.0; du[i] = 0.0; }
  const double beta = P.beta[p];
  double Z = 0.0;
  {
    uint32_t mw[W];
#pragma unroll
    for (int w = 0; w < W; w++) mw[w] = P.mask[(size_t)w * np + p];
    int wcur = 0, node[NLPS_GROUP];
    while (next_group<W>(m, base, mw, wcur, node) > 0) {
      double Xn[NLPS_GROUP][D], An[NLPS_GROUP][D], dUn[NLPS_GROUP][D];
#pragma unroll
      for (int u = 0; u < NLPS_GROUP; u++)
#pragma unroll
        for (int i = 0; i < D; i++) { Xn[u][i] = 0.0; An[u][i] = 0.0; dUn[u][i] = 0.0; }
#pragma unroll
      for (int u = 0; u < NLPS_GROUP; u++)
        if (node[u] >= 0) {
          ldvec<D>(&m.X[(size_t)node[u] * NS<D>::X], Xn[u]);
          ldvec<D>(&G.UA[(size_t)node[u] * 2 * NS<D>::X], dUn[u]);
          ldvec<D>(&G.UA[(size_t)node[u] * 2 * NS<D>::X + NS<D>::X], An[u]);
        }
#pragma unroll
      for (int u = 0; u < NLPS_GROUP; u++) {
        if (node[u] < 0) continue;
        double ll = 0.0, lx = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) {
          double l = xp[i] - Xn[u][i];
          ll += l * l;
          lx += l * lam[i];
        }
        double e = exp(-beta * ll + lx);
        Z += e;
#pragma unroll
        for (int i = 0; i < D; i++) {
          a[i] += e * An[u][i];
          du[i] += e * dUn[u][i];
        }
      }
    }
  }
  const double Zi = 1.0 / Z;
#pragma unroll
  for (int i = 0; i < D; i++) {
    double ai = a[i] * Zi, di = du[i] * Zi;
    P.acc[i * np + p] = ai;
    P.ddis[i * np + p] = di;
    P.vel[i * np + p] += 0.5 * sp.dt * ai;
    P.x[i * np + p] = xp[i] + di;
    P.dis[i * np + p] += di;
  }
}


// Reference roll semantics for history variables of particles whose law never writes them
// (Neo-Hookean: b_e, EPS, Kappa): U-Verlet.c:1043-1056 COPIES n1 -> n every step, so after the
// first step both copies hold the initial n1 value.  With the pointer-swap roll that is reproduced by
// copying n1 -> n once, before the first swap.
template <int D>
__global__ void k_sync_inert(PartDev P) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.np) return;
  if (c_mat[P.matidx[p]].type != NLPS_MAT_NEO_HOOKEAN_WRIGGERS) return;
  constexpr int TB = (D == 2) ? 5 : 9;
#pragma unroll
  for (int i = 0; i < TB; i++) P.be_n[(size_t)i * P.np + p] = P.be_n1[(size_t)i * P.np + p];
  P.eps_n[p] = P.eps_n1[p];
  P.kap_n[p] = P.kap_n1[p];
}

// AoS (host layout, n x cols) <-> SoA (cols x n)
__global__ void k_aos_to_soa(const double* aos, double* soa, int n, int cols, int aos_stride, int col0) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * cols) return;
  int c = (int)(i / n), p = (int)(i % n);
  soa[i] = aos[(size_t)p * aos_stride + col0 + c];
}
__global__ void k_soa_to_aos(const double* soa, double* aos, int n, int cols, int aos_stride, int col0) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * cols) return;
  int p = (int)(i / cols), c = (int)(i % cols);
  aos[(size_t)p * aos_stride + col0 + c] = soa[(size_t)c * n + p];
}
__global__ void k_fill_d(double* a, size_t n, double v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}
// expand bitmask lists to the reference's ListNodes order (reverse of acceptance order)
__global__ void k_expand_lists(MeshDev m, PartDev P, int W, int cap, int* lists) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.np) return;
  int base = m.r2p[P.I0[p]], o = 0;
  for (int w = W - 1; w >= 0; w--) {
    uint32_t mm = P.mask[(size_t)w * P.np + p];
    while (mm) {
      int b = 31 - __clz(mm);
      mm &= ~(1u << b);
      if (o < cap) lists[(size_t)p * cap + o] = m.r2i[base + w * 32 + b];
      o++;
    }
  }
  for (; o < cap; o++) lists[(size_t)p * cap + o] = -1;
}
__global__ void k_export_nodal(GridDev G, int nn, int D, int which, double* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)nn * D) return;
  int A = (int)(i / D), k = (int)(i % D);
  const int xs = (D == 2) ? 2 : 4;
  double v = 0.0;
  if (G.active[A]) {
    bool fx = (G.fixed[A] >> k) & 1u;
    switch (which) {
      case 0: v = G.M[A]; break;
      case 1: v = G.UA[(size_t)A * 2 * xs + k]; break;
      case 2: v = G.F[i]; break;
      case 3: v = G.UA[(size_t)A * 2 * xs + xs + k]; break;
      case 4: v = fx ? G.F[i] : 0.0; break;
    }
  }
  out[i] = v;
}


// Stress_integration__Constitutive__ (Constitutive.c:18-258) on arrays of material points
// (AoS host layout), the GPU twin used by the point-wise parity tests.
template <int D>
__global__ void k_stress_points(int n, int mat, ReturnMapParams rp, const double* DF, const double* F1, const double* J1,
                                const double* be_n, const double* eps_n, const double* kap_n, double* stress,
                                double* be_n1, double* eps_n1, double* kap_n1, double* W, double* cep, int* status) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  constexpr int T = (D == 2) ? 5 : 9;
  const MatParams& m = c_mat[mat];
  double df[D * D], f1[D * D], be[T], tau[T], c[D * D], Wp = 0.0;
#pragma unroll
  for (int i = 0; i < D * D; i++) { df[i] = DF[(size_t)p * T + i]; f1[i] = F1[(size_t)p * T + i]; c[i] = 0.0; }
#pragma unroll
  for (int i = 0; i < T; i++) be[i] = be_n[(size_t)p * T + i];
  double eps = eps_n[p], kap = kap_n[p];
  int st = 0;
  if (m.type == NLPS_MAT_NEO_HOOKEAN_WRIGGERS) stress_neo_hookean<D>(m, f1, J1[p], tau, Wp);
  else if (m.type == NLPS_MAT_DRUCKER_PRAGER) st = stress_drucker_prager<D>(m, rp, df, be, eps, kap, tau, Wp, c);
  else st = stress_matsuoka_nakai<D>(m, rp, df, be, eps, kap, tau, Wp, c);
  status[p] = st;
#pragma unroll
  for (int i = 0; i < T; i++) { stress[(size_t)p * T + i] = tau[i]; be_n1[(size_t)p * T + i] = be[i]; }
  eps_n1[p] = eps; kap_n1[p] = kap; W[p] = Wp;
#pragma unroll
  for (int i = 0; i < D * D; i++) cep[(size_t)p * D * D + i] = c[i];
}

// ---------------------------------------------------------------------------
// Host-side engine
struct nlps_engine {
  int D = 2, T = 5, TB = 5, W = 1, np = 0, nn = 0, device = 0, cap = 0;
  nlps_solver solver{};
  cudaStream_t stream = nullptr;
  MeshDev mesh{};
  PartDev P{};
  GridDev G{};
  BcDev bc{};
  NeuDev neu{};
  double* grav = nullptr;
  int* err = nullptr;
  int* h_err = nullptr;  // pinned
  int* npart_check = nullptr;
  double neg_log_tol = 0.0, dt = 0.0;
  int has_traction = 0;
  int max_occ = 0, max_act = 0;
  int cap_r2 = 0;  // largest 2-ring row
  int uniform_mat = -1;  // material type shared by all materials, or -1
  int inert_synced = 0;
  std::vector<void*> allocs;
  // staging for AoS <-> SoA
  double* stage = nullptr;
  size_t stage_doubles = 0;
  double* h_stage = nullptr;  // pinned
  // profiling
  int profile = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double k_ms[K_COUNT] = {0};
  int k_n[K_COUNT] = {0};
  long long launches = 0;
  int last_code = 0, last_particle = -1;
};

template <typename Tp>
static int dev_alloc(nlps_engine* e, Tp** p, size_t n) {
  void* q = nullptr;
  cudaError_t st = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(Tp));
  if (st != cudaSuccess) {
    fprintf(stderr, "nlps_b200: cudaMalloc(%zu) failed: %s\n", n * sizeof(Tp), cudaGetErrorString(st));
    return 1;
  }
  cudaMemsetAsync(q, 0, std::max<size_t>(n, 1) * sizeof(Tp), e->stream);
  e->allocs.push_back(q);
  *p = (Tp*)q;
  return 0;
}
template <typename Tp>
static int dev_upload(nlps_engine* e, Tp** p, const Tp* h, size_t n) {
  if (dev_alloc(e, p, n)) return 1;
  if (n) CUDA_OK(cudaMemcpyAsync(*p, h, n * sizeof(Tp), cudaMemcpyHostToDevice, e->stream));
  return 0;
}

static void transpose_csr(int nn, const int* ptr, const int* idx, std::vector<int>& tp, std::vector<int>& ti,
                          std::vector<unsigned char>* qpos = nullptr) {
  tp.assign(nn + 1, 0);
  for (int i = 0; i < nn; i++)
    for (int q = ptr[i]; q < ptr[i + 1]; q++) tp[idx[q] + 1]++;
  for (int i = 0; i < nn; i++) tp[i + 1] += tp[i];
  ti.resize(tp[nn]);
  if (qpos) qpos->resize(ptr[nn]);
  std::vector<int> fill(tp.begin(), tp.end() - 1);
  for (int i = 0; i < nn; i++)
    for (int q = ptr[i]; q < ptr[i + 1]; q++) {
      int A = idx[q];
      if (qpos) (*qpos)[q] = (unsigned char)(fill[A] - tp[A]);
      ti[fill[A]++] = i;
    }
}

#define LAUNCH(e, id, kernel, grid, block, ...)                                 \
  do {                                                                          \
    if ((e)->profile) cudaEventRecord((e)->ev0, (e)->stream);                   \
    kernel<<<(grid), (block), 0, (e)->stream>>>(__VA_ARGS__);                   \
    (e)->launches++;                                                            \
    if ((e)->profile) {                                                         \
      cudaEventRecord((e)->ev1, (e)->stream);                                   \
      cudaEventSynchronize((e)->ev1);                                           \
      float _ms = 0;                                                            \
      cudaEventElapsedTime(&_ms, (e)->ev0, (e)->ev1);                           \
      (e)->k_ms[id] += _ms;                                                     \
      (e)->k_n[id]++;                                                           \
    }                                                                           \
  } while (0)

static inline int nblk(size_t n, int b) { return (int)((n + b - 1) / b); }

static StepParams make_params(nlps_engine* e, int step, int update_I0) {
  StepParams sp;
  sp.dt = e->dt;
  sp.gamma_lme = e->solver.gamma_lme;
  sp.neg_log_tol = e->neg_log_tol;
  sp.tol_wrapper = e->solver.tol_wrapper_lme;
  sp.thickness = e->solver.thickness;
  sp.max_iter_lme = e->solver.max_iter_lme;
  sp.nsteps = e->solver.num_steps;
  sp.step = step;
  sp.update_I0 = update_I0;
  sp.W = e->W;
  sp.rp.tol = e->solver.tol_radial_returning;
  sp.rp.max_iter = e->solver.max_iter_radial_returning;
  sp.rp.quirk_rows = e->solver.quirk_transposed_eigvec;
  sp.rp.want_cep = e->solver.compute_c_ep;
  return sp;
}

// AoS host -> SoA device for one field (cols columns starting at col0 of an aos_stride-wide row)
static int put_field(nlps_engine* e, const double* h, double* d, int cols, int aos_stride, int col0) {
  if (!h || !d) return 0;
  size_t n = (size_t)e->np * aos_stride;
  CUDA_OK(cudaMemcpyAsync(e->stage, h, n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  k_aos_to_soa<<<nblk((size_t)e->np * cols, 256), 256, 0, e->stream>>>(e->stage, d, e->np, cols, aos_stride, col0);
  CUDA_OK(cudaStreamSynchronize(e->stream));  // host buffer may be pageable; stage is reused
  return 0;
}
static int get_field(nlps_engine* e, double* h, const double* d, int cols, int aos_stride, int col0,
                     const double* d_extra = nullptr) {
  if (!h || !d) return 0;
  size_t n = (size_t)e->np * aos_stride;
  if (cols != aos_stride) {
    // partial rows (2D tensors: 4 in-plane + slot 4): assemble the whole row on the device
    if (d_extra) k_soa_to_aos<<<nblk((size_t)e->np, 256), 256, 0, e->stream>>>(d_extra, e->stage, e->np, 1, aos_stride, cols);
  }
  k_soa_to_aos<<<nblk((size_t)e->np * cols, 256), 256, 0, e->stream>>>(d, e->stage, e->np, cols, aos_stride, col0);
  CUDA_OK(cudaMemcpyAsync(h, e->stage, n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CUDA_OK(cudaStreamSynchronize(e->stream));
  return 0;
}

static int poll_error(nlps_engine* e) {
  if (cudaMemcpyAsync(e->h_err, e->err, 2 * sizeof(int), cudaMemcpyDeviceToHost, e->stream) != cudaSuccess ||
      cudaStreamSynchronize(e->stream) != cudaSuccess) {
    cudaError_t st = cudaGetLastError();
    fprintf(stderr, "nlps_b200: CUDA failure while polling: %s\n", cudaGetErrorString(st));
    e->last_code = NLPS_ERR_CUDA;
    return 1;
  }
  cudaError_t st = cudaGetLastError();
  if (st != cudaSuccess) {
    fprintf(stderr, "nlps_b200: kernel launch failure: %s\n", cudaGetErrorString(st));
    e->last_code = NLPS_ERR_CUDA;
    return 1;
  }
  if (e->h_err[0] != 0) {
    e->last_code = e->h_err[0];
    e->last_particle = e->h_err[1];
    // reference convention: RED message on stderr naming the failing routine (U-Verlet.c:101-135)
    static const char* what[] = {"", "", "I3__TensorLib__(F_n1_p): negative jacobian", "tributary__LME__: insufficient nodal connectivity",
                                 "__lambda_Newton_Rapson: Hessian near to singular matrix", "__lambda_Newton_Rapson: no convergence",
                                 "compute_Kirchhoff_Stress_Drucker_Prager__Constitutive__", "compute_Kirchhoff_Stress_Matsuoka_Nakai__Constitutive__",
                                 "compute_adjunt__TensorLib__: singular DF"};
    fprintf(stderr, "\033[31mError in %s (particle %d)\033[0m\n", e->h_err[0] < 9 ? what[e->h_err[0]] : "device", e->h_err[1]);
    return 1;
  }
  return 0;
}

template <int D>
static void stage_search_t(nlps_engine* e, int step, int update_I0, int do_predictor) {
  const int np = e->np, nn = e->nn;
  cudaMemsetAsync(e->G.cnt, 0, sizeof(int) * nn, e->stream);
  LAUNCH(e, K_SEARCH, k_search<D>, nblk(np, 256), 256, e->mesh, e->P, e->G, update_I0);
  LAUNCH(e, K_NODE_FLAGS, k_node_flags, nblk(nn, 256), 256, e->mesh, e->G);
  int nb = nblk(nn, SCAN_ITEMS);
  LAUNCH(e, K_SCAN1, k_scan_reduce, nb, 256, e->G.packed, e->G.scan_blk, nn);
  LAUNCH(e, K_SCAN2, k_scan_tops, 1, 1024, e->G.scan_blk, nb, e->G.n_active, e->G.n_occ, e->npart_check);
  LAUNCH(e, K_SCAN3, k_scan_apply, nb, 256, e->G.packed, e->G.scan_blk, e->G.cell_start, e->G.occ_pos, e->G.act_pos, nn);
  LAUNCH(e, K_FILL, k_cell_fill, nblk(np, 256), 256, e->P, e->G);
  LAUNCH(e, K_NODE_FINISH, k_node_finish, nblk(nn, 256), 256, e->mesh, e->G);
  StepParams sp = make_params(e, step, update_I0);
  switch (e->W) {
#define CASE_W(w) case w: { auto kfn = k_lme<D, w>; LAUNCH(e, K_LME, kfn, nblk(np, 128), 128, e->mesh, e->P, e->G, sp, e->err, do_predictor); } break;
    CASE_W(1) CASE_W(2) CASE_W(4) CASE_W(8)
#undef CASE_W
  }
}
template <int D, bool FORCE>
static void launch_p2g_cell(nlps_engine* e, int id) {
  const int it = (e->cap_r2 + 31) / 32;
  const int grid = nblk((size_t)e->max_occ, 4 * (it == 1 ? 4 : 1));  // 4 warps per block, CPW cells per warp
  switch (it) {
    case 1: { auto kfn = k_p2g_cell<D, 1, FORCE>; LAUNCH(e, id, kfn, grid, 128, e->mesh, e->P, e->G, e->has_traction); } break;
    case 2: { auto kfn = k_p2g_cell<D, 2, FORCE>; LAUNCH(e, id, kfn, grid, 128, e->mesh, e->P, e->G, e->has_traction); } break;
    case 3: case 4: { auto kfn = k_p2g_cell<D, 4, FORCE>; LAUNCH(e, id, kfn, grid, 128, e->mesh, e->P, e->G, e->has_traction); } break;
    default: { auto kfn = k_p2g_cell<D, 8, FORCE>; LAUNCH(e, id, kfn, grid, 128, e->mesh, e->P, e->G, e->has_traction); } break;
  }
}
template <int D>
static void stage_p2g_mass_disp_t(nlps_engine* e, int step) {
  // at most min(nn, np) cells are occupied; one warp per cell (surplus warps exit on n_occ)
  launch_p2g_cell<D, false>(e, K_P2G_MASS_DISP);
  LAUNCH(e, K_GRID_DISP, k_grid_disp<D>, nblk(e->max_act, 128), 128, e->mesh, e->G, e->bc, step);
}
template <int D>
static void stage_kin_stress_t(nlps_engine* e, int step) {
  StepParams sp = make_params(e, step, 1);
  switch (e->W) {
#define CASE_WM(w, mt) { auto kfn = k_kin_stress<D, w, mt>; LAUNCH(e, K_KIN_STRESS, kfn, nblk(e->np, 128), 128, e->mesh, e->P, e->G, sp, e->err); }
#define CASE_W(w) case w: switch (e->uniform_mat) { case 0: CASE_WM(w, 0) break; case 1: CASE_WM(w, 1) break; case 2: CASE_WM(w, 2) break; default: CASE_WM(w, -1) break; } break;
    CASE_W(1) CASE_W(2) CASE_W(4) CASE_W(8)
#undef CASE_W
#undef CASE_WM
  }
}
template <int D>
static void stage_force_t(nlps_engine* e, int step) {
  if (e->has_traction) {
    LAUNCH(e, K_TRACTION, k_traction_clear<D>, nblk(e->np, 256), 256, e->P);
    LAUNCH(e, K_TRACTION, k_traction<D>, nblk(e->neu.n_entries, 128), 128, e->P, e->neu, e->solver.thickness, step);
  }
  launch_p2g_cell<D, true>(e, K_P2G_FORCE);
  LAUNCH(e, K_GRID_ACC, k_grid_acc<D>, nblk(e->max_act, 128), 128, e->mesh, e->G, e->grav, e->solver.num_steps, step);
}
template <int D>
static void stage_g2p_t(nlps_engine* e, int step) {
  StepParams sp = make_params(e, step, 1);
  switch (e->W) {
#define CASE_W(w) case w: { auto kfn = k_g2p<D, w>; LAUNCH(e, K_G2P, kfn, nblk(e->np, 128), 128, e->mesh, e->P, e->G, sp); } break;
    CASE_W(1) CASE_W(2) CASE_W(4) CASE_W(8)
#undef CASE_W
  }
  if (!e->inert_synced) {
    k_sync_inert<D><<<nblk(e->np, 256), 256, 0, e->stream>>>(e->P);
    e->inert_synced = 1;
  }
  // roll n+1 -> n (U-Verlet.c:1043-1081) as pointer swaps
  std::swap(e->P.F_n, e->P.F_n1);
  std::swap(e->P.J_n, e->P.J_n1);
  std::swap(e->P.be_n, e->P.be_n1);
  std::swap(e->P.eps_n, e->P.eps_n1);
  std::swap(e->P.kap_n, e->P.kap_n1);
}

static void enqueue_stage(nlps_engine* e, int stage, int step) {
  const bool d2 = e->D == 2;
  switch (stage) {
    case NLPS_STAGE_SEARCH: d2 ? stage_search_t<2>(e, step, 1, 1) : stage_search_t<3>(e, step, 1, 1); break;
    case NLPS_STAGE_P2G_MASS_DISP: d2 ? stage_p2g_mass_disp_t<2>(e, step) : stage_p2g_mass_disp_t<3>(e, step); break;
    case NLPS_STAGE_KIN_STRESS: d2 ? stage_kin_stress_t<2>(e, step) : stage_kin_stress_t<3>(e, step); break;
    case NLPS_STAGE_FORCE: d2 ? stage_force_t<2>(e, step) : stage_force_t<3>(e, step); break;
    case NLPS_STAGE_G2P: d2 ? stage_g2p_t<2>(e, step) : stage_g2p_t<3>(e, step); break;
    default: break;  // GRID_DISP / GRID_ACC are fused into the node kernels
  }
}

// ---------------------------------------------------------------------------
extern "C" {

const char* nlps_b200_version(void) { return "nlps_b200 0.1 (sm_100a, fp64, explicit NPC-FS)"; }

void nlps_b200_destroy(nlps_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (void* p : e->allocs) cudaFree(p);
  if (e->h_err) cudaFreeHost(e->h_err);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

static int set_err(char* err, int len, const char* msg) {
  if (err && len > 0) { strncpy(err, msg, len - 1); err[len - 1] = 0; }
  fprintf(stderr, "nlps_b200_create: %s\n", msg);
  return 1;
}

static int create_impl(nlps_engine* e, const nlps_mesh* mesh, const nlps_solver* solver, int n_bounds,
                       const nlps_load* bounds, int n_neumann, const nlps_load* neumann, const double* gravity,
                       int n_materials, const nlps_material* materials, const nlps_particles* st, char* err,
                       int err_len) {
  const int D = mesh->ndim, nn = mesh->n_nodes, np = st->n;
  e->D = D; e->T = (D == 2) ? 5 : 9; e->TB = e->T; e->nn = nn; e->np = np;
  e->solver = *solver;
  if (e->solver.quirk_transposed_eigvec < 0) e->solver.quirk_transposed_eigvec = (D == 2) ? 1 : 0;
  if (D != 2 && D != 3) return set_err(err, err_len, "ndim must be 2 or 3");
  if (n_materials < 1 || n_materials > MAX_MATERIALS) return set_err(err, err_len, "1..8 materials supported");
  if (!st->x_GC || !st->mass || !st->Vol_0 || !st->rho || !st->I0 || !st->MatIdx)
    return set_err(err, err_len, "x_GC, mass, Vol_0, rho, I0 and MatIdx are mandatory");
  CUDA_OK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreate(&e->ev0));
  CUDA_OK(cudaEventCreate(&e->ev1));
  CUDA_OK(cudaMallocHost(&e->h_err, 2 * sizeof(int)));
  e->dt = solver->cfl * mesh->delta_x / solver->cel;  // Courant.c:6-55 (DynamicTimeStep = false)
  e->neg_log_tol = -log(solver->tol_zero_lme);        // host libm, LME.c:1052
  // ---- mesh
  int maxr2 = 0;
  for (int i = 0; i < nn; i++) maxr2 = std::max(maxr2, mesh->ring2_ptr[i + 1] - mesh->ring2_ptr[i]);
  e->cap = maxr2;
  e->cap_r2 = maxr2;
  e->W = (maxr2 + 31) / 32;
  while (e->W & (e->W - 1)) e->W++;  // kernels are instantiated for 1, 2, 4, 8 mask words
  if (e->W > MAX_MASK_WORDS) return set_err(err, err_len, "2-ring larger than 256 nodes is not supported");
  double* dX; int *r1p, *r1i, *r2p, *r2i, *t1p, *t1i, *t2p, *t2i; double* dh;
  {
    const int xs = (D == 2) ? 2 : 4;
    std::vector<double> Xp((size_t)nn * xs, 0.0);
    for (int i = 0; i < nn; i++)
      for (int k = 0; k < D; k++) Xp[(size_t)i * xs + k] = mesh->coords[(size_t)i * D + k];
    if (dev_upload(e, &dX, Xp.data(), Xp.size())) return 1;
    CUDA_OK(cudaStreamSynchronize(e->stream));
  }
  if (dev_upload(e, &r1p, mesh->ring1_ptr, (size_t)nn + 1)) return 1;
  if (dev_upload(e, &r1i, mesh->ring1_idx, (size_t)mesh->ring1_ptr[nn])) return 1;
  if (dev_upload(e, &r2p, mesh->ring2_ptr, (size_t)nn + 1)) return 1;
  if (dev_upload(e, &r2i, mesh->ring2_idx, (size_t)mesh->ring2_ptr[nn])) return 1;
  if (dev_upload(e, &dh, mesh->h_avg, (size_t)nn)) return 1;
  std::vector<int> tp, ti;
  transpose_csr(nn, mesh->ring1_ptr, mesh->ring1_idx, tp, ti);
  if (dev_upload(e, &t1p, tp.data(), tp.size())) return 1;
  if (dev_upload(e, &t1i, ti.data(), ti.size())) return 1;
  CUDA_OK(cudaStreamSynchronize(e->stream));
  std::vector<unsigned char> qpos;
  transpose_csr(nn, mesh->ring2_ptr, mesh->ring2_idx, tp, ti, &qpos);
  int maxr2t = 0, maxr1 = 0;
  for (int i = 0; i < nn; i++) {
    maxr2t = std::max(maxr2t, tp[i + 1] - tp[i]);
    maxr1 = std::max(maxr1, mesh->ring1_ptr[i + 1] - mesh->ring1_ptr[i]);
  }
  if (maxr2t > 255) return set_err(err, err_len, "transposed 2-ring larger than 255 nodes is not supported");
  unsigned char* dq;
  if (dev_upload(e, &t2p, tp.data(), tp.size())) return 1;
  if (dev_upload(e, &t2i, ti.data(), ti.size())) return 1;
  if (dev_upload(e, &dq, qpos.data(), qpos.size())) return 1;
  CUDA_OK(cudaStreamSynchronize(e->stream));
  e->mesh = MeshDev{nn, dX, r1p, r1i, r2p, r2i, t1p, t1i, t2p, t2i, dq, dh};
  e->max_occ = (int)std::min<long long>(nn, np);
  e->max_act = (int)std::min<long long>(nn, (long long)np * maxr1);
  // ---- grid work arrays
  GridDev& G = e->G;
  if (dev_alloc(e, &G.M, nn) || dev_alloc(e, &G.UA, (size_t)nn * 2 * (D == 2 ? 2 : 4)) || dev_alloc(e, &G.F, (size_t)nn * D) ||
      dev_alloc(e, &G.active, nn) || dev_alloc(e, &G.fixed, nn) ||
      dev_alloc(e, &G.cnt, nn) || dev_alloc(e, &G.cursor, nn) || dev_alloc(e, &G.cell_start, nn) ||
      dev_alloc(e, &G.plist, np) || dev_alloc(e, &G.act_list, nn) || dev_alloc(e, &G.n_active, 1) ||
      dev_alloc(e, &G.packed, nn) || dev_alloc(e, &G.scan_blk, (size_t)nblk(nn, SCAN_ITEMS) + 1) ||
      dev_alloc(e, &G.act_pos, nn) || dev_alloc(e, &G.occ_pos, nn) || dev_alloc(e, &G.occ_list, e->max_occ) ||
      dev_alloc(e, &G.n_occ, 1) || dev_alloc(e, &e->npart_check, 1) || dev_alloc(e, &e->err, 2))
    return 1;
  G.cap = maxr2t;
  if (dev_alloc(e, &G.part, (size_t)e->max_act * G.cap * (1 + D))) return 1;
  // ---- boundary conditions: node -> boundaries CSR (boundary order preserved)
  {
    int maxdim = 1;
    for (int b = 0; b < n_bounds; b++) maxdim = std::max(maxdim, bounds[b].dim);
    std::vector<int> np_(nn + 1, 0), dims(std::max(n_bounds, 1), 0);
    for (int b = 0; b < n_bounds; b++) {
      dims[b] = bounds[b].dim;
      for (int j = 0; j < bounds[b].n_ids; j++) {
        int A = bounds[b].ids[j];
        if (A < 0 || A >= nn) return set_err(err, err_len, "Dirichlet node id out of range");
        np_[A + 1]++;
      }
    }
    for (int i = 0; i < nn; i++) np_[i + 1] += np_[i];
    std::vector<int> nb(std::max(np_[nn], 1)), fill(np_.begin(), np_.end() - 1);
    for (int b = 0; b < n_bounds; b++)
      for (int j = 0; j < bounds[b].n_ids; j++) nb[fill[bounds[b].ids[j]]++] = b;
    const int ns = solver->num_steps;
    std::vector<int> dir((size_t)std::max(n_bounds, 1) * maxdim * ns, 0);
    std::vector<double> val((size_t)std::max(n_bounds, 1) * maxdim * ns, 0.0);
    for (int b = 0; b < n_bounds; b++)
      for (int k = 0; k < bounds[b].dim; k++)
        for (int s = 0; s < ns; s++) {
          dir[((size_t)b * maxdim + k) * ns + s] = bounds[b].dir[(size_t)k * ns + s];
          val[((size_t)b * maxdim + k) * ns + s] = bounds[b].val[(size_t)k * ns + s];
        }
    int *d_np, *d_nb, *d_dims, *d_dir; double* d_val;
    if (dev_upload(e, &d_np, np_.data(), np_.size()) || dev_upload(e, &d_nb, nb.data(), nb.size()) ||
        dev_upload(e, &d_dims, dims.data(), dims.size()) || dev_upload(e, &d_dir, dir.data(), dir.size()) ||
        dev_upload(e, &d_val, val.data(), val.size()))
      return 1;
    CUDA_OK(cudaStreamSynchronize(e->stream));
    e->bc = BcDev{d_np, d_nb, d_dims, d_dir, d_val, maxdim, ns, n_bounds};
  }
  // ---- Neumann loads
  e->has_traction = 0;
  if (n_neumann > 0) {
    int maxdim = 1, tot = 0;
    for (int b = 0; b < n_neumann; b++) { maxdim = std::max(maxdim, neumann[b].dim); tot += neumann[b].n_ids; }
    std::vector<int> part(std::max(tot, 1)), load(std::max(tot, 1)), dims(n_neumann);
    int o = 0;
    for (int b = 0; b < n_neumann; b++) {
      dims[b] = neumann[b].dim;
      for (int j = 0; j < neumann[b].n_ids; j++) {
        if (neumann[b].ids[j] < 0 || neumann[b].ids[j] >= np) return set_err(err, err_len, "Neumann particle id out of range");
        part[o] = neumann[b].ids[j]; load[o] = b; o++;
      }
    }
    const int ns = solver->num_steps;
    std::vector<int> dir((size_t)n_neumann * maxdim * ns, 0);
    std::vector<double> val((size_t)n_neumann * maxdim * ns, 0.0);
    for (int b = 0; b < n_neumann; b++)
      for (int k = 0; k < neumann[b].dim; k++)
        for (int s = 0; s < ns; s++) {
          dir[((size_t)b * maxdim + k) * ns + s] = neumann[b].dir[(size_t)k * ns + s];
          val[((size_t)b * maxdim + k) * ns + s] = neumann[b].val[(size_t)k * ns + s];
        }
    int *d_part, *d_load, *d_dims, *d_dir; double* d_val;
    if (dev_upload(e, &d_part, part.data(), part.size()) || dev_upload(e, &d_load, load.data(), load.size()) ||
        dev_upload(e, &d_dims, dims.data(), dims.size()) || dev_upload(e, &d_dir, dir.data(), dir.size()) ||
        dev_upload(e, &d_val, val.data(), val.size()))
      return 1;
    CUDA_OK(cudaStreamSynchronize(e->stream));
    e->neu = NeuDev{tot, d_part, d_load, d_dims, d_dir, d_val, maxdim, ns};
    e->has_traction = tot > 0;
  }
  if (gravity) {
    if (dev_upload(e, &e->grav, gravity, (size_t)D * solver->num_steps)) return 1;
    CUDA_OK(cudaStreamSynchronize(e->stream));
  }
  // ---- materials
  {
    MatParams hm[MAX_MATERIALS];
    memset(hm, 0, sizeof(hm));
    for (int i = 0; i < n_materials; i++) {
      const nlps_material& s = materials[i];
      if (s.type < 0 || s.type > 2) return set_err(err, err_len, "unknown material type");
      hm[i] = MatParams{s.type, s.rho, s.E, s.nu, s.reference_pressure, s.kappa_0, s.hardening_modulus,
                        s.plastic_strain_0, s.phi_frictional, s.psi_frictional, s.exponent_hardening_ortiz,
                        s.cohesion, s.alpha_hardening_borja, s.a_hardening_borja[0], s.a_hardening_borja[1],
                        s.a_hardening_borja[2]};
    }
    CUDA_OK(cudaMemcpyToSymbol(c_mat, hm, sizeof(hm)));
    e->uniform_mat = hm[0].type;
    for (int i = 1; i < n_materials; i++)
      if (hm[i].type != hm[0].type) e->uniform_mat = -1;
  }
  // ---- particles
  PartDev& P = e->P;
  P.np = np;
  const int T = e->T, DD = D * D, TBv = (D == 2) ? 5 : 9;
#define A_(f, c) if (dev_alloc(e, &P.f, (size_t)np * (c))) return 1;
  A_(x, D) A_(dis, D) A_(ddis, D) A_(vel, D) A_(acc, D) A_(lam, D)
  A_(beta, 1) A_(mass, 1) A_(vol0, 1) A_(rho, 1) A_(W, 1)
  A_(J_n, 1) A_(J_n1, 1) A_(eps_n, 1) A_(eps_n1, 1) A_(kap_n, 1) A_(kap_n1, 1)
  A_(F_n, DD) A_(F_n1, DD) A_(DF, DD) A_(be_n, TBv) A_(be_n1, TBv) A_(stress, T) A_(cep, DD)
  A_(Fs4, 1) A_(DFs4, 1)
#undef A_
  if (dev_alloc(e, &P.rec, (size_t)np * (D == 2 ? Rec<2>::SIZE : Rec<3>::SIZE))) return 1;
  if (dev_alloc(e, &P.I0, np) || dev_alloc(e, &P.nnodes, np) || dev_alloc(e, &P.matidx, np) ||
      dev_alloc(e, &P.mask, (size_t)np * e->W))
    return 1;
  e->stage_doubles = (size_t)np * std::max(T, DD);
  if (dev_alloc(e, &e->stage, e->stage_doubles)) return 1;
  // defaults as allocate_U_vars__Fields__ leaves them (identity tensors, J = 1)
  auto fill = [&](double* a, size_t n, double v) { k_fill_d<<<nblk(n, 256), 256, 0, e->stream>>>(a, n, v); };
  for (int i = 0; i < D; i++) {
    fill(P.F_n + (size_t)(i * D + i) * np, np, 1.0);
    fill(P.F_n1 + (size_t)(i * D + i) * np, np, 1.0);
    fill(P.DF + (size_t)(i * D + i) * np, np, 1.0);
    fill(P.be_n + (size_t)(i * D + i) * np, np, 1.0);
    fill(P.be_n1 + (size_t)(i * D + i) * np, np, 1.0);
  }
  if (D == 2) { fill(P.be_n + (size_t)4 * np, np, 1.0); fill(P.be_n1 + (size_t)4 * np, np, 1.0); }
  fill(P.Fs4, np, 1.0); fill(P.DFs4, np, 1.0);
  fill(P.J_n, np, 1.0); fill(P.J_n1, np, 1.0);
  CUDA_OK(cudaStreamSynchronize(e->stream));
  if (nlps_b200_upload(e, st)) return 1;
  CUDA_OK(cudaMemcpyAsync(P.I0, st->I0, sizeof(int) * np, cudaMemcpyHostToDevice, e->stream));
  CUDA_OK(cudaMemcpyAsync(P.matidx, st->MatIdx, sizeof(int) * np, cudaMemcpyHostToDevice, e->stream));
  CUDA_OK(cudaStreamSynchronize(e->stream));
  for (int p = 0; p < np; p++)
    if (st->MatIdx[p] < 0 || st->MatIdx[p] >= n_materials) return set_err(err, err_len, "MatIdx out of range");
  for (int p = 0; p < np; p++)
    if (st->I0[p] < 0 || st->I0[p] >= nn) return set_err(err, err_len, "I0 out of range");
  return 0;
}

nlps_engine* nlps_b200_create(const nlps_mesh* mesh, const nlps_solver* solver, int n_bounds, const nlps_load* bounds,
                              int n_neumann, const nlps_load* neumann, const double* gravity, int n_materials,
                              const nlps_material* materials, const nlps_particles* state, int device, char* err,
                              int err_len) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_err(err, err_len, "no CUDA device: this library has no CPU fallback");
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    set_err(err, err_len, "cudaSetDevice failed");
    return nullptr;
  }
  nlps_engine* e = new nlps_engine();
  e->device = device;
  if (create_impl(e, mesh, solver, n_bounds, bounds, n_neumann, neumann, gravity, n_materials, materials, state, err,
                  err_len)) {
    nlps_b200_destroy(e);
    return nullptr;
  }
  return e;
}

int nlps_b200_upload(nlps_engine* e, const nlps_particles* in) {
  cudaSetDevice(e->device);
  if (in->b_e_n || in->b_e_n1 || in->EPS_n || in->EPS_n1 || in->Kappa_n || in->Kappa_n1) e->inert_synced = 0;
  const int D = e->D, T = e->T, DD = D * D;
  PartDev& P = e->P;
  if (put_field(e, in->x_GC, P.x, D, D, 0) || put_field(e, in->dis, P.dis, D, D, 0) ||
      put_field(e, in->D_dis, P.ddis, D, D, 0) || put_field(e, in->vel, P.vel, D, D, 0) ||
      put_field(e, in->acc, P.acc, D, D, 0) || put_field(e, in->lambda, P.lam, D, D, 0))
    return 1;
  if (put_field(e, in->F_n, P.F_n, DD, T, 0) || put_field(e, in->F_n1, P.F_n1, DD, T, 0) ||
      put_field(e, in->DF, P.DF, DD, T, 0))
    return 1;
  if (D == 2) {
    if (put_field(e, in->F_n, P.Fs4, 1, T, 4) || put_field(e, in->DF, P.DFs4, 1, T, 4)) return 1;
  }
  if (put_field(e, in->b_e_n, P.be_n, T, T, 0) || put_field(e, in->b_e_n1, P.be_n1, T, T, 0) ||
      put_field(e, in->Stress, P.stress, T, T, 0) || put_field(e, in->C_ep, P.cep, DD, DD, 0))
    return 1;
  if (put_field(e, in->J_n, P.J_n, 1, 1, 0) || put_field(e, in->J_n1, P.J_n1, 1, 1, 0) ||
      put_field(e, in->mass, P.mass, 1, 1, 0) || put_field(e, in->rho, P.rho, 1, 1, 0) ||
      put_field(e, in->Vol_0, P.vol0, 1, 1, 0) || put_field(e, in->W, P.W, 1, 1, 0) ||
      put_field(e, in->EPS_n, P.eps_n, 1, 1, 0) || put_field(e, in->EPS_n1, P.eps_n1, 1, 1, 0) ||
      put_field(e, in->Kappa_n, P.kap_n, 1, 1, 0) || put_field(e, in->Kappa_n1, P.kap_n1, 1, 1, 0) ||
      put_field(e, in->Beta, P.beta, 1, 1, 0))
    return 1;
  return 0;
}

int nlps_b200_download(nlps_engine* e, nlps_particles* out) {
  cudaSetDevice(e->device);
  const int D = e->D, T = e->T, DD = D * D;
  PartDev& P = e->P;
  if (get_field(e, out->x_GC, P.x, D, D, 0) || get_field(e, out->dis, P.dis, D, D, 0) ||
      get_field(e, out->D_dis, P.ddis, D, D, 0) || get_field(e, out->vel, P.vel, D, D, 0) ||
      get_field(e, out->acc, P.acc, D, D, 0) || get_field(e, out->lambda, P.lam, D, D, 0))
    return 1;
  if (D == 2) {
    if (get_field(e, out->F_n, P.F_n, DD, T, 0, P.Fs4) || get_field(e, out->F_n1, P.F_n1, DD, T, 0, P.Fs4) ||
        get_field(e, out->DF, P.DF, DD, T, 0, P.DFs4))
      return 1;
  } else {
    if (get_field(e, out->F_n, P.F_n, DD, T, 0) || get_field(e, out->F_n1, P.F_n1, DD, T, 0) ||
        get_field(e, out->DF, P.DF, DD, T, 0))
      return 1;
  }
  if (get_field(e, out->b_e_n, P.be_n, T, T, 0) || get_field(e, out->b_e_n1, P.be_n1, T, T, 0) ||
      get_field(e, out->Stress, P.stress, T, T, 0) || get_field(e, out->C_ep, P.cep, DD, DD, 0))
    return 1;
  if (get_field(e, out->J_n, P.J_n, 1, 1, 0) || get_field(e, out->J_n1, P.J_n1, 1, 1, 0) ||
      get_field(e, out->mass, P.mass, 1, 1, 0) || get_field(e, out->rho, P.rho, 1, 1, 0) ||
      get_field(e, out->Vol_0, P.vol0, 1, 1, 0) || get_field(e, out->W, P.W, 1, 1, 0) ||
      get_field(e, out->EPS_n, P.eps_n, 1, 1, 0) || get_field(e, out->EPS_n1, P.eps_n1, 1, 1, 0) ||
      get_field(e, out->Kappa_n, P.kap_n, 1, 1, 0) || get_field(e, out->Kappa_n1, P.kap_n1, 1, 1, 0) ||
      get_field(e, out->Beta, P.beta, 1, 1, 0))
    return 1;
  if (out->I0) CUDA_OK(cudaMemcpyAsync(out->I0, P.I0, sizeof(int) * e->np, cudaMemcpyDeviceToHost, e->stream));
  if (out->NumberNodes) CUDA_OK(cudaMemcpyAsync(out->NumberNodes, P.nnodes, sizeof(int) * e->np, cudaMemcpyDeviceToHost, e->stream));
  CUDA_OK(cudaStreamSynchronize(e->stream));
  return 0;
}

int nlps_b200_initialize_lme(nlps_engine* e) {
  cudaSetDevice(e->device);
  // Beta as given (0 after allocation => infinite radius, Appendix D-4); I0 is NOT moved.
  if (e->D == 2) stage_search_t<2>(e, 0, 0, 0); else stage_search_t<3>(e, 0, 0, 0);
  return poll_error(e);
}

int nlps_b200_stage(nlps_engine* e, int stage, int time_step) {
  cudaSetDevice(e->device);
  enqueue_stage(e, stage, time_step);
  return poll_error(e);
}

int nlps_b200_run(nlps_engine* e, int first_step, int count) {
  cudaSetDevice(e->device);
  for (int k = first_step; k < first_step + count; k++)
    for (int s = NLPS_STAGE_SEARCH; s <= NLPS_STAGE_G2P; s++) enqueue_stage(e, s, k);
  return poll_error(e);
}

int nlps_b200_step(nlps_engine* e, int time_step) { return nlps_b200_run(e, time_step, 1); }

int nlps_b200_timed_run(nlps_engine* e, int first_step, int count, double* ms) {
  cudaSetDevice(e->device);
  cudaEvent_t a, b;
  CUDA_OK(cudaEventCreate(&a));
  CUDA_OK(cudaEventCreate(&b));
  CUDA_OK(cudaStreamSynchronize(e->stream));
  CUDA_OK(cudaEventRecord(a, e->stream));
  for (int k = first_step; k < first_step + count; k++)
    for (int s = NLPS_STAGE_SEARCH; s <= NLPS_STAGE_G2P; s++) enqueue_stage(e, s, k);
  CUDA_OK(cudaEventRecord(b, e->stream));
  CUDA_OK(cudaEventSynchronize(b));
  float t = 0;
  CUDA_OK(cudaEventElapsedTime(&t, a, b));
  *ms = t;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return poll_error(e);
}

int nlps_b200_get_nodal(nlps_engine* e, int which, double* out) {
  cudaSetDevice(e->device);
  if (which < 0 || which > 4) return 1;
  size_t n = (size_t)e->nn * e->D;
  double* tmp = nullptr;
  CUDA_OK(cudaMalloc(&tmp, n * sizeof(double)));
  k_export_nodal<<<nblk(n, 256), 256, 0, e->stream>>>(e->G, e->nn, e->D, which, tmp);
  cudaError_t st = cudaMemcpyAsync(out, tmp, n * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
  if (st == cudaSuccess) st = cudaStreamSynchronize(e->stream);
  cudaFree(tmp);
  return st == cudaSuccess ? 0 : 1;
}

int nlps_b200_get_active(nlps_engine* e, unsigned char* out) {
  cudaSetDevice(e->device);
  CUDA_OK(cudaMemcpyAsync(out, e->G.active, e->nn, cudaMemcpyDeviceToHost, e->stream));
  CUDA_OK(cudaStreamSynchronize(e->stream));
  return 0;
}

int nlps_b200_list_capacity(nlps_engine* e) { return e->cap; }

int nlps_b200_get_lists(nlps_engine* e, int* counts, int* lists, int cap) {
  cudaSetDevice(e->device);
  int* tmp = nullptr;
  size_t n = (size_t)e->np * cap;
  CUDA_OK(cudaMalloc(&tmp, n * sizeof(int)));
  k_expand_lists<<<nblk(e->np, 128), 128, 0, e->stream>>>(e->mesh, e->P, e->W, cap, tmp);
  cudaError_t st = cudaMemcpyAsync(lists, tmp, n * sizeof(int), cudaMemcpyDeviceToHost, e->stream);
  if (st == cudaSuccess && counts)
    st = cudaMemcpyAsync(counts, e->P.nnodes, sizeof(int) * e->np, cudaMemcpyDeviceToHost, e->stream);
  if (st == cudaSuccess) st = cudaStreamSynchronize(e->stream);
  cudaFree(tmp);
  return st == cudaSuccess ? 0 : 1;
}

int nlps_b200_last_error(nlps_engine* e, int* code, int* particle) {
  if (code) *code = e->last_code;
  if (particle) *particle = e->last_particle;
  return e->last_code;
}

double nlps_b200_dt(nlps_engine* e) { return e->dt; }

int nlps_b200_profile(nlps_engine* e, int enable) { e->profile = enable; return 0; }
int nlps_b200_kernel_times(nlps_engine* e, int cap, const char** names, double* ms, int* launches) {
  for (int i = 0; i < K_COUNT && i < cap; i++) {
    if (names) names[i] = kKernelNames[i];
    if (ms) ms[i] = e->k_ms[i];
    if (launches) launches[i] = e->k_n[i];
  }
  return K_COUNT;
}
void nlps_b200_reset_kernel_times(nlps_engine* e) {
  for (int i = 0; i < K_COUNT; i++) { e->k_ms[i] = 0; e->k_n[i] = 0; }
}
long long nlps_b200_launch_count(nlps_engine* e) { return e->launches; }

int nlps_b200_u_verlet(const nlps_mesh* mesh, const nlps_solver* solver, int n_bounds, const nlps_load* bounds,
                       int n_neumann, const nlps_load* neumann, const double* gravity, int n_materials,
                       const nlps_material* materials, nlps_particles* state, int run_initialize, int results_every,
                       nlps_results_cb cb, void* user, int device) {
  char msg[256];
  nlps_engine* e = nlps_b200_create(mesh, solver, n_bounds, bounds, n_neumann, neumann, gravity, n_materials, materials,
                                    state, device, msg, sizeof(msg));
  if (!e) return 1;
  int status = 0;
  if (run_initialize) status = nlps_b200_initialize_lme(e);
  int k = solver->initial_step;
  while (!status && k < solver->num_steps) {
    int chunk = solver->num_steps - k;
    if (results_every > 0) {  // results after every step with TimeStep % ResultsTimeStep == 0 (U-Verlet.c:1097)
      int nxt = ((k + results_every - 1) / results_every) * results_every;
      chunk = std::min(chunk, nxt - k + 1);
    }
    status = nlps_b200_run(e, k, chunk);
    k += chunk;
    if (!status && results_every > 0 && ((k - 1) % results_every == 0)) {
      status = nlps_b200_download(e, state);
      if (!status && cb) cb(k - 1, user);
    }
  }
  if (!status) status = nlps_b200_download(e, state);
  nlps_b200_destroy(e);
  return status;
}

int nlps_b200_stress_points(int ndim, const nlps_material* material, double tol_radial, int max_iter_radial,
                            int quirk_transposed_eigvec, int n, const double* DF, const double* F_n1,
                            const double* J_n1, const double* b_e_n, const double* eps_n, const double* kappa_n,
                            double* stress, double* b_e_n1, double* eps_n1, double* kappa_n1, double* W,
                            double* C_ep, int* status, int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "nlps_b200_stress_points: no CUDA device: this library has no CPU fallback\n");
    return 1;
  }
  CUDA_OK(cudaSetDevice(device));
  const int T = ndim == 2 ? 5 : 9, DD = ndim * ndim;
  const nlps_material& sm = *material;
  MatParams hm[MAX_MATERIALS];
  memset(hm, 0, sizeof(hm));
  hm[0] = MatParams{sm.type, sm.rho, sm.E, sm.nu, sm.reference_pressure, sm.kappa_0, sm.hardening_modulus,
                    sm.plastic_strain_0, sm.phi_frictional, sm.psi_frictional, sm.exponent_hardening_ortiz,
                    sm.cohesion, sm.alpha_hardening_borja, sm.a_hardening_borja[0], sm.a_hardening_borja[1],
                    sm.a_hardening_borja[2]};
  CUDA_OK(cudaMemcpyToSymbol(c_mat, hm, sizeof(hm)));
  ReturnMapParams rp{tol_radial, max_iter_radial, quirk_transposed_eigvec < 0 ? (ndim == 2) : quirk_transposed_eigvec, 1};
  size_t nT = (size_t)n * T;
  double* d = nullptr;
  int* dst = nullptr;
  size_t tot = nT * 5 + (size_t)n * 6 + (size_t)n * DD;
  CUDA_OK(cudaMalloc(&d, tot * sizeof(double)));
  CUDA_OK(cudaMalloc(&dst, n * sizeof(int)));
  double *dDF = d, *dF1 = dDF + nT, *dbe = dF1 + nT, *dS = dbe + nT, *dbe1 = dS + nT, *dJ = dbe1 + nT, *deps = dJ + n,
         *dkap = deps + n, *deps1 = dkap + n, *dkap1 = deps1 + n, *dW = dkap1 + n, *dC = dW + n;
  cudaMemcpy(dDF, DF, nT * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dF1, F_n1, nT * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dbe, b_e_n, nT * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dJ, J_n1, n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(deps, eps_n, n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dkap, kappa_n, n * 8, cudaMemcpyHostToDevice);
  if (ndim == 2) k_stress_points<2><<<nblk(n, 64), 64>>>(n, 0, rp, dDF, dF1, dJ, dbe, deps, dkap, dS, dbe1, deps1, dkap1, dW, dC, dst);
  else k_stress_points<3><<<nblk(n, 64), 64>>>(n, 0, rp, dDF, dF1, dJ, dbe, deps, dkap, dS, dbe1, deps1, dkap1, dW, dC, dst);
  cudaError_t st = cudaDeviceSynchronize();
  if (st == cudaSuccess) {
    cudaMemcpy(stress, dS, nT * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(b_e_n1, dbe1, nT * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(eps_n1, deps1, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(kappa_n1, dkap1, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(W, dW, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(C_ep, dC, (size_t)n * DD * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(status, dst, n * sizeof(int), cudaMemcpyDeviceToHost);
  } else {
    fprintf(stderr, "nlps_b200_stress_points: %s\n", cudaGetErrorString(st));
  }
  cudaFree(d);
  cudaFree(dst);
  return st == cudaSuccess ? 0 : 1;
}

}  // extern "C"
