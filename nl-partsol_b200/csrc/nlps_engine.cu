// nlps_engine.cu -- B200-native (sm_100a) engine for NL-PartSol's explicit NPC-FS step.
//
// Design: fp64 everywhere, particle state SoA in HBM, mesh adjacency as
// CSR in the reference's chain order.  Neighbour lists are stored as BITMASKS over the
// 2-ring of the closest node (4 B / 16 B per particle instead of 4n B).  Particles are binned by
// closest node (I0) every step and physically re-sorted into that order every few steps.  One
// thread block owns a run of consecutive occupied cells: it stages the 2-ring node data of its
// cells in shared memory once, runs the per-particle work out of shared memory (no dependent
// global gathers), and then assembles the cell's nodal sums with a warp per cell and a lane per
// node from the shape-function weights still in shared memory: ATOMICS-FREE particle-to-grid
// (shared-memory fp64 atomicAdd is a CAS loop on sm_100a), deterministic summation order.  Grid
// update + Dirichlet BCs are fused into the per-node reduction kernels; state roll is a pointer swap.
//
// Reference citations are relative to nl-partsol/src of migmolper/NL-PartSol.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is bound at run time (see nccl_api)
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <string>
#include <mutex>
#include <unistd.h>
#include <vector>

#include "../../include/nlps_b200.h"
#include "nlps_device.cuh"
#include "nlps_types.cuh"
#include "nlps_cellwarp.h"

#define CUDA_OK(call)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      fprintf(stderr, "nlps_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)


enum KernelId {
  K_SEARCH = 0, K_MARK, K_NODE_FLAGS, K_SCAN1, K_SCAN2, K_SCAN3, K_FILL, K_NODE_FINISH, K_REORDER, K_LME_P2G,
  K_GRID_DISP, K_TRACTION, K_KIN_FORCE, K_GRID_ACC, K_G2P, K_HALO, K_KIN_GATHER, K_STRESS, K_COUNT
};
static const char* kKernelNames[K_COUNT] = {
    "search_closest_node", "mark_live_blocks", "node_flags", "scan_reduce", "scan_tops", "scan_apply", "cell_fill",
    "node_finish", "reorder", "lme_p2g_mass_disp", "grid_disp_bc", "traction", "kin_stress_p2g_force",
    "grid_acc", "g2p_update", "halo_exchange", "kin_gather", "stress_update"};

struct Carve {
  size_t off = 0;
  __host__ __device__ size_t take(size_t bytes) { size_t o = off; off = (off + bytes + 15) & ~(size_t)15; return o; }
};
// shared-memory layouts, evaluated identically on host (size) and device (offsets)
template <int D, int W, bool CACHE>
struct LayoutA {  // k_lme_p2g
  size_t tab, cs, base, len, B, cs2, base2, len2, B2, rank, q, X, pa, zinv, mass, ddis, mask, px, plam, pbeta, cw, pre, ovf, total;
  __host__ __device__ LayoutA(const BlockCfg& c) {
    Carve k;
    const size_t pairs = (size_t)c.C * c.SL, pc = c.PCAP;
    tab = k.take(8 * 32);
    cs = k.take(4 * (c.C + 1)); base = k.take(4 * c.C); len = k.take(4 * c.C); B = k.take(4 * c.C);
    cs2 = k.take(4 * (c.C + 1)); base2 = k.take(4 * c.C); len2 = k.take(4 * c.C); B2 = k.take(4 * c.C);  // next group (pipeline)
    rank = k.take(4 * pairs); q = k.take(pairs); X = k.take(8 * D * pairs);
    pa = CACHE ? k.take(8 * pc * c.SL) : 0;
    zinv = k.take(8 * pc); mass = k.take(8 * pc); ddis = k.take(8 * D * pc); mask = k.take(4 * W * pc);
    px = plam = pbeta = 0;
    if (!CACHE) { px = k.take(8 * D * pc); plam = k.take(8 * D * pc); pbeta = k.take(8 * pc); }
    // compact weight cache (3D): the weights of a particle's neighbours in ascending slot order + the prefix
    // popcounts of its mask words, so that the cell phase finds weight(j, k) without another exp
    cw = pre = 0;
    ovf = k.take(16);
    if (!CACHE && c.NCA > 0) { cw = k.take(8 * pc * c.NCA); pre = k.take(pc * W); }
    total = k.off;
  }
};
template <int D, int W, bool CACHE>
struct LayoutB {  // k_kin_force
  size_t tab, cs, base, len, B, cs2, base2, len2, B2, rank, q, X, U, pa, zinv, G, px, trac, mask, plam, pbeta, total;
  __host__ __device__ LayoutB(const BlockCfg& c) {
    Carve k;
    const size_t pairs = (size_t)c.C * c.SL, pc = c.PCAP;
    tab = k.take(8 * 32);
    cs = k.take(4 * (c.C + 1)); base = k.take(4 * c.C); len = k.take(4 * c.C); B = k.take(4 * c.C);
    cs2 = k.take(4 * (c.C + 1)); base2 = k.take(4 * c.C); len2 = k.take(4 * c.C); B2 = k.take(4 * c.C);  // next group (pipeline)
    rank = k.take(4 * pairs); q = k.take(pairs); X = k.take(8 * D * pairs); U = k.take(8 * D * pairs);
    pa = CACHE ? k.take(8 * pc * c.SL) : 0;
    zinv = k.take(8 * pc); G = k.take(8 * D * D * pc); px = k.take(8 * D * pc); trac = k.take(8 * D * pc);
    mask = k.take(4 * W * pc);
    plam = pbeta = 0;
    if (!CACHE) { plam = k.take(8 * D * pc); pbeta = k.take(8 * pc); }
    total = k.off;
  }
};
template <int D>
struct LayoutC {  // k_g2p
  size_t tab, cs, base, len, B, cs2, base2, len2, B2, rank, X, U, A, total;
  __host__ __device__ LayoutC(const BlockCfg& c) {
    Carve k;
    const size_t pairs = (size_t)c.C * c.SL;
    tab = k.take(8 * 32);
    cs = k.take(4 * (c.C + 1)); base = k.take(4 * c.C); len = k.take(4 * c.C); B = k.take(4 * c.C);
    cs2 = k.take(4 * (c.C + 1)); base2 = k.take(4 * c.C); len2 = k.take(4 * c.C); B2 = k.take(4 * c.C);  // next group (pipeline)
    rank = k.take(4 * pairs); X = k.take(8 * D * pairs); U = k.take(8 * D * pairs); A = k.take(8 * D * pairs);
    total = k.off;
  }
};


// s* of a search radius (see sstar_from_Ra): per node for beta = gamma / h_avg^2 (beta__LME__, LME.c:177-185), and
// per particle for the beta it carries (the PREVIOUS step's, LME.c:973,983-984; 0 after allocation => infinite radius)
__global__ void __launch_bounds__(256) k_sstar_nodes(const double* h_avg, int nn, double gamma_lme, double neg_log_tol, double* sst) {
  const int A = blockIdx.x * blockDim.x + threadIdx.x;
  if (A >= nn) return;
  const double h = h_avg[A];
  const double beta = __ddiv_rn(gamma_lme, __dmul_rn(h, h));
  sst[A] = sstar_from_Ra(__dsqrt_rn(__ddiv_rn(neg_log_tol, beta)));
}
__global__ void __launch_bounds__(256) k_sstar_particles(const double* beta, int np, double neg_log_tol, double* sstar) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= np) return;
  sstar[p] = sstar_from_Ra(__dsqrt_rn(__ddiv_rn(neg_log_tol, beta[p])));
}

// ---------------------------------------------------------------------------
// K0a: closest node + cell histogram.   local_search__LME__ first loop (LME.c:917-944),
// get_closest_node__MeshTools__ (Nodes-Tools.c:476-538): first strict minimum, chain order.
template <int D>
__global__ void __launch_bounds__(256) k_search(MeshDev m, PartDev P, GridDev G, int update_I0, SlabDev sl, int* err) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.np) return;
  int I0 = P.I0[p];
  if (update_I0) {
    double nd = 0.0, xp[D];
#pragma unroll
    for (int i = 0; i < D; i++) {
      double dd = P.dis[i * P.ld + p];
      nd = __dadd_rn(nd, __dmul_rn(dd, dd));
      xp[i] = P.x[i * P.ld + p];
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
      if (best != I0) P.I0[p] = best;
      I0 = best;
    }
  }
  if (sl.on) {
    const double c = m.X[(size_t)I0 * NS<D>::X + sl.axis];
    if (c < sl.lo || c > sl.hi) latch_error(err, NLPS_ERR_SLAB_EXCURSION, P.orig[p]);
  }
  atomicAdd(&G.cnt[I0], 1);
  G.occ_blk[I0 >> 8] = 1;
}

// Only the node blocks within the 2-ring of an occupied cell can change: the node kernels below skip the rest
// (the benchmark grid is 6x wider than the column; 4/5 of its nodes never see a particle).
__global__ void __launch_bounds__(256) k_mark(MeshDev m, GridDev G) {
  if (!G.occ_blk[blockIdx.x]) return;
  const int A = blockIdx.x * 256 + threadIdx.x;
  if (A >= m.nn) return;
  if (G.cnt[A] > 0 || G.rocc[A]) {
    // test before set: thousands of threads mark the same few bytes (the stale read is a benign race)
    // (consecutive ring entries mostly fall into the same 256-node block: one flag access per run; the 1-ring is a
    // subset of the 2-ring wherever both exist, but nothing here relies on it)
    int last = -1;
    for (int q = m.r2p[A]; q < m.r2p[A + 1]; q++) { const int b = m.r2i[q] >> 8; if (b != last) { last = b; if (!G.dirty_cur[b]) G.dirty_cur[b] = 1; } }
    for (int q = m.r1p[A]; q < m.r1p[A + 1]; q++) { const int b = m.r1i[q] >> 8; if (b != last) { last = b; if (!G.dirty_cur[b]) G.dirty_cur[b] = 1; } }
    if ((A >> 8) != last && !G.dirty_cur[A >> 8]) G.dirty_cur[A >> 8] = 1;
  }
}
__global__ void __launch_bounds__(256) k_live(GridDev G, int nblocks) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  G.live[b] = G.dirty_cur[b] | G.dirty_prev[b];
  G.dirty_prev[b] = 0;  // becomes dirty_cur of the next step
  G.occ_blk[b] = 0;
}

// node flags: ActiveNode[A] = OR over particles with I0 in {B : A in ring1(B)} (LME.c:949-965,
// after the reset of Shape-Functions.c:38-47).  packed.x = cnt | occupied << 40, packed.y = active:
// one fused exclusive scan yields cell_start, the rank of each occupied cell and of each active node.
__global__ void __launch_bounds__(256) k_node_flags(MeshDev m, GridDev G) {
  if (!G.live[blockIdx.x]) return;
  int A = blockIdx.x * blockDim.x + threadIdx.x;
  if (A >= m.nn) return;
  int act = 0;
  for (int q = m.r1tp[A]; q < m.r1tp[A + 1] && !act; q++) {
    const int B = m.r1ti[q];
    act = G.cnt[B] > 0 || G.rocc[B];
  }
  G.active[A] = (unsigned char)act;
  unsigned long long c = (unsigned)G.cnt[A];
  G.packed[A] = make_ulonglong2(c | ((unsigned long long)(c > 0) << 40), (unsigned long long)act);
  G.cursor[A] = 0;
}

__device__ __forceinline__ ulonglong2 add2(ulonglong2 a, ulonglong2 b) { return make_ulonglong2(a.x + b.x, a.y + b.y); }

// exclusive scan of packed, 3 phases, 2048 items per block
static const int SCAN_ITEMS = 2048;
// live (optional): one flag per 256 items; a chunk of SCAN_ITEMS whose flags are all clear holds only zeros
__device__ __forceinline__ bool chunk_live(const unsigned char* live, int chunk, int n) {
  if (!live) return true;
  const int b0 = chunk * (SCAN_ITEMS / 256), nb = (n + 255) / 256;
  unsigned v = 0;
  for (int k = 0; k < SCAN_ITEMS / 256 && b0 + k < nb; k++) v |= live[b0 + k];
  return v != 0;
}
__global__ void __launch_bounds__(256) k_scan_reduce(const ulonglong2* in, ulonglong2* blk, int n, const unsigned char* live = nullptr) {
  __shared__ ulonglong2 sh[256];
  if (!chunk_live(live, blockIdx.x, n)) {
    if (threadIdx.x == 0) blk[blockIdx.x] = make_ulonglong2(0, 0);
    return;
  }
  size_t base = (size_t)blockIdx.x * SCAN_ITEMS;
  ulonglong2 s = make_ulonglong2(0, 0);
  for (int i = threadIdx.x; i < SCAN_ITEMS; i += 256)
    if (base + i < (size_t)n && (!live || live[(base + i) >> 8])) s = add2(s, in[base + i]);
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
                                                    int* occ_pos, int* act_pos, int n, const unsigned char* live = nullptr) {
  __shared__ ulonglong2 sh[256];
  if (!chunk_live(live, blockIdx.x, n)) return;  // nobody reads the positions of nodes outside the live blocks
  size_t base = (size_t)blockIdx.x * SCAN_ITEMS;
  const int per = SCAN_ITEMS / 256;
  ulonglong2 v[per], s = make_ulonglong2(0, 0);
  // a thread's `per` items lie in one block of 256 (per divides 256): clean blocks are neither read nor written
  const bool on = !live || live[(base + (size_t)threadIdx.x * per) >> 8];
#pragma unroll
  for (int k = 0; k < per; k++) {
    size_t i = base + (size_t)threadIdx.x * per + k;
    v[k] = (on && i < (size_t)n) ? in[i] : make_ulonglong2(0, 0);
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
    if (on && i < (size_t)n) {
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

// per node: sort the cell's particles by the CALLER's particle id (summation order of the cell sums is
// then independent of the physical particle order), append occupied cells / active nodes to their
// compact lists, and record which slots of the node's transposed 2-ring belong to occupied cells.
__global__ void __launch_bounds__(256) k_node_finish(MeshDev m, PartDev P, GridDev G) {
  if (!G.live[blockIdx.x]) return;
  int A = blockIdx.x * blockDim.x + threadIdx.x;
  if (A >= m.nn) return;
  int n = G.cnt[A];
  if (n > 1) {
    int* a = G.plist + G.cell_start[A];
    for (int i = 1; i < n; i++) {
      int v = a[i], kv = P.orig[v], j = i - 1;
      while (j >= 0 && P.orig[a[j]] > kv) { a[j + 1] = a[j]; j--; }
      a[j + 1] = v;
    }
  }
  if (n > 0) {
    G.occ_list[G.occ_pos[A]] = A;
    const int bs = m.r2p[A];
    // .w = 2-ring length (9 bits) | particles of the cell << 9
    G.occ_meta[G.occ_pos[A]] = make_int4(A, G.cell_start[A], bs, (m.r2p[A + 1] - bs) | (n << 9));
  }
  int rank = -1;
  if (G.active[A]) {
    rank = G.act_pos[A];
    G.act_list[rank] = A;
    const int q0 = m.r2tp[A], nq = m.r2tp[A + 1] - q0;
    for (int w = 0; w < G.w2t; w++) {
      uint32_t word = 0u;
      for (int b = 0; b < 32; b++) {
        int q = w * 32 + b;
        if (q < nq && G.cnt[m.r2ti[q0 + q]] > 0) word |= 1u << b;
      }
      G.occm[(size_t)w * G.max_act + rank] = word;
    }
  }
  G.arank[A] = rank;
}

// physical re-sort: gather every SoA field into the cell-sorted order (dst[c][t] = src[c][plist[t]])
template <typename Tp>
__global__ void __launch_bounds__(256) k_gather_rows(const Tp* src, Tp* dst, const int* plist, int np, int ld, int cols) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)np * cols) return;
  int c = (int)(i / np), t = (int)(i % np);
  dst[(size_t)c * ld + t] = src[(size_t)c * ld + plist[t]];
}
__global__ void __launch_bounds__(256) k_after_sort(PartDev P, GridDev G) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.np) return;
  G.plist[t] = t;
  P.inv[P.orig[t]] = t;
}

// ---------------------------------------------------------------------------
// block prologue shared by the three cell-block kernels
#ifndef NLPS_STAGE_U
#define NLPS_STAGE_U 4
#endif
#ifndef NLPS_STAGE_U_G2P
#define NLPS_STAGE_U_G2P 4
#endif
struct Blk { int ncell, t0, t1; };
// metadata of cell group g (C consecutive occupied cells): one global round trip
__device__ __forceinline__ void blk_prologue(const GridDev& G, const BlockCfg& cfg, int nocc, int np, int g, int* s_cs,
                                             int* s_base, int* s_len, int* s_B, Blk& b) {
  const int c0 = g * cfg.C;
  b.ncell = min(cfg.C, nocc - c0);
  for (int i = threadIdx.x; i <= b.ncell; i += blockDim.x) {
    if (c0 + i < nocc) {
      const int4 mt = G.occ_meta[c0 + i];
      s_cs[i] = mt.y;
      if (i < b.ncell) { s_B[i] = mt.x; s_base[i] = mt.z; s_len[i] = mt.w & 511; }
    } else {
      s_cs[i] = np;  // the last occupied cell ends at the last particle
    }
  }
  __syncthreads();
  b.t0 = s_cs[0];
  b.t1 = s_cs[b.ncell];
}
// stage the 2-ring node data of the block's cells: every (cell, slot) pair is loaded once, four pairs per
// thread in flight (ids first, then all their data: two global round trips for the whole block); the
// per-particle loops then run out of shared memory with no dependent global gathers.
template <int D, bool WANT_Q, int NF>
__device__ __forceinline__ void stage_nodes(const MeshDev& m, const GridDev& G, int SL, unsigned magic, int ncell,
                                            const int* s_base, const int* s_len, int* s_rank, unsigned char* s_q,
                                            double* s_X, double* s_U, double* s_A) {
  // pairs in flight per thread (7 = one sweep of the usual block, measured no faster than 4 and costs registers)
  constexpr int U = (NF == 2) ? NLPS_STAGE_U_G2P : NLPS_STAGE_U;
  const int npairs = ncell * SL;
  for (int e0 = threadIdx.x; e0 < npairs; e0 += U * blockDim.x) {
    int node[U], idx[U], es[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int e = e0 + u * blockDim.x;
      node[u] = -1;
      idx[u] = 0;
      es[u] = -1;
      if (e < npairs) {
        const int c = (int)(((unsigned)e * magic) >> 21), k = e - c * SL;
        es[u] = e;
        if (k < s_len[c]) { idx[u] = s_base[c] + k; node[u] = m.r2i[idx[u]]; }
      }
    }
    // invalid pairs load node 0 (harmless) so that every load below is unconditional and independent
    int rank[U];
    unsigned char qv[U];
    double2 x0[U], x1[U], u0[U], u1[U], a0[U], a1[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int nd = max(node[u], 0);
      const double* px = &m.X[(size_t)nd * NS<D>::X];
      const double* pu = &G.UA[(size_t)nd * 2 * NS<D>::X];
      rank[u] = G.arank[nd];
      x0[u] = *reinterpret_cast<const double2*>(px);
      if (D == 3) x1[u] = *reinterpret_cast<const double2*>(px + 2);
      if (WANT_Q) qv[u] = G.cm_sl ? m.r2pos[idx[u]] : m.r2q[idx[u]];  // cell-major sums: position in the cell's run
      if (NF >= 1) {
        u0[u] = *reinterpret_cast<const double2*>(pu);
        if (D == 3) u1[u] = *reinterpret_cast<const double2*>(pu + 2);
      }
      if (NF >= 2) {
        a0[u] = *reinterpret_cast<const double2*>(pu + NS<D>::X);
        if (D == 3) a1[u] = *reinterpret_cast<const double2*>(pu + NS<D>::X + 2);
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int e = es[u];
      if (e >= 0) {
        s_rank[e] = (node[u] >= 0) ? rank[u] : -1;
        if (WANT_Q) s_q[e] = qv[u];
        double* dx = s_X + (size_t)e * D;
        dx[0] = x0[u].x; dx[1] = x0[u].y;
        if (D == 3) dx[2] = x1[u].x;
        if (NF >= 1) {
          double* du = s_U + (size_t)e * D;
          du[0] = u0[u].x; du[1] = u0[u].y;
          if (D == 3) du[2] = u1[u].x;
        }
        if (NF >= 2) {
          double* da = s_A + (size_t)e * D;
          da[0] = a0[u].x; da[1] = a0[u].y;
          if (D == 3) da[2] = a1[u].x;
        }
      }
    }
  }
}
// ---- cell-group pipeline -----------------------------------------------------------------------------------------
// A persistent block walks over the groups g, g + grid, g + 2 grid, ...  While it works on group g, the metadata of
// group g + 2 grid and the ring node ids of group g + grid are already in flight (registers), so that the staging of
// the next group starts with its ids at hand: one dependent global round trip (node data) instead of three
// (metadata -> ring ids -> node data).
constexpr int NLPS_NID = 8;  // ring ids a thread holds for the next group: C * SL <= NLPS_NID * threads (checked at create)
struct MetaPtrs { int *cs, *base, *len, *B; };
// thread i <-> cell c0 + i of group g (i <= C): the record travels through a register
__device__ __forceinline__ int4 meta_fetch(const GridDev& G, const BlockCfg& cfg, int nocc, int g) {
  int4 mt = make_int4(0, 0, 0, 0);
  const int i = threadIdx.x, c0 = g * cfg.C;
  if (i <= cfg.C && c0 + i < nocc) mt = G.occ_meta[c0 + i];
  return mt;
}
__device__ __forceinline__ void meta_put(const BlockCfg& cfg, int nocc, int np, int g, const int4& mt, const MetaPtrs& s) {
  const int i = threadIdx.x, c0 = g * cfg.C, ncell = min(cfg.C, nocc - c0);
  if (i <= ncell) {
    if (c0 + i < nocc) {
      s.cs[i] = mt.y;
      if (i < ncell) { s.B[i] = mt.x; s.base[i] = mt.z; s.len[i] = mt.w & 511; }
    } else {
      s.cs[i] = np;  // the last occupied cell ends at the last particle
    }
  }
}
__device__ __forceinline__ void ids_fetch(const MeshDev& m, const BlockCfg& cfg, int ncell, const MetaPtrs& s, int (&ids)[NLPS_NID]) {
  const int npairs = ncell * cfg.SL;
#pragma unroll
  for (int u = 0; u < NLPS_NID; u++) {
    const int e = threadIdx.x + u * blockDim.x;
    ids[u] = -1;
    if (e < npairs) {
      const int c = (int)(((unsigned)e * cfg.magic) >> 21), k = e - c * cfg.SL;
      if (k < s.len[c]) ids[u] = m.r2i[s.base[c] + k];
    }
  }
}
// stage_nodes with the ring ids already in registers (pair e = threadIdx.x + u * blockDim.x <-> ids[u])
template <int D, bool WANT_Q, int NF>
__device__ __forceinline__ void stage_nodes_ids(const MeshDev& m, const GridDev& G, const BlockCfg& cfg, int ncell,
                                                const MetaPtrs& s, const int (&ids)[NLPS_NID], int* s_rank,
                                                unsigned char* s_q, double* s_X, double* s_U, double* s_A) {
  constexpr int U = 4;
  static_assert(NLPS_NID % U == 0, "batches");
  const int npairs = ncell * cfg.SL, SL = cfg.SL;
#pragma unroll
  for (int b0 = 0; b0 < NLPS_NID; b0 += U) {
    if ((int)(b0 * blockDim.x) >= npairs) break;
    int rank[U];
    unsigned char qv[U];
    double2 x0[U], x1[U], u0[U], u1[U], a0[U], a1[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int nd = max(ids[b0 + u], 0);  // invalid pairs load node 0 (harmless): unconditional independent loads
      const double* px = &m.X[(size_t)nd * NS<D>::X];
      const double* pu = &G.UA[(size_t)nd * 2 * NS<D>::X];
      rank[u] = G.arank[nd];
      x0[u] = *reinterpret_cast<const double2*>(px);
      if (D == 3) x1[u] = *reinterpret_cast<const double2*>(px + 2);
      if (WANT_Q) {
        const int e = threadIdx.x + (b0 + u) * blockDim.x;
        qv[u] = 0;
        if (ids[b0 + u] >= 0) {
          const int c = (int)(((unsigned)e * cfg.magic) >> 21), k = e - c * SL;
          qv[u] = G.cm_sl ? m.r2pos[s.base[c] + k] : m.r2q[s.base[c] + k];
        }
      }
      if (NF >= 1) {
        u0[u] = *reinterpret_cast<const double2*>(pu);
        if (D == 3) u1[u] = *reinterpret_cast<const double2*>(pu + 2);
      }
      if (NF >= 2) {
        a0[u] = *reinterpret_cast<const double2*>(pu + NS<D>::X);
        if (D == 3) a1[u] = *reinterpret_cast<const double2*>(pu + NS<D>::X + 2);
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int e = threadIdx.x + (b0 + u) * blockDim.x;
      if (e < npairs) {
        s_rank[e] = (ids[b0 + u] >= 0) ? rank[u] : -1;
        if (WANT_Q) s_q[e] = qv[u];
        double* dx = s_X + (size_t)e * D;
        dx[0] = x0[u].x; dx[1] = x0[u].y;
        if (D == 3) dx[2] = x1[u].x;
        if (NF >= 1) {
          double* du = s_U + (size_t)e * D;
          du[0] = u0[u].x; du[1] = u0[u].y;
          if (D == 3) du[2] = u1[u].x;
        }
        if (NF >= 2) {
          double* da = s_A + (size_t)e * D;
          da[0] = a0[u].x; da[1] = a0[u].y;
          if (D == 3) da[2] = a1[u].x;
        }
      }
    }
  }
}
// loop head / tail of the pipelined walk (used by the three cell-block kernels)
#define NLPS_PIPE_BEGIN(WANT_Q, NF, SQ, SU, SA)                                                                        \
  MetaPtrs mc{(int*)(smem + L.cs), (int*)(smem + L.base), (int*)(smem + L.len), (int*)(smem + L.B)};                    \
  MetaPtrs mn{(int*)(smem + L.cs2), (int*)(smem + L.base2), (int*)(smem + L.len2), (int*)(smem + L.B2)};                \
  int ids[NLPS_NID];                                                                                                   \
  {                                                                                                                    \
    const int g0 = blockIdx.x, g1 = blockIdx.x + gridDim.x;                                                            \
    if (g0 < ngroups) meta_put(cfg, nocc, P.np, g0, meta_fetch(G, cfg, nocc, g0), mc);                                  \
    if (g1 < ngroups) meta_put(cfg, nocc, P.np, g1, meta_fetch(G, cfg, nocc, g1), mn);                                  \
    __syncthreads();                                                                                                   \
    if (g0 < ngroups) ids_fetch(m, cfg, min(cfg.C, nocc - g0 * cfg.C), mc, ids);                                        \
  }                                                                                                                    \
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {                                                        \
    int* const s_cs = mc.cs; int* const s_base = mc.base; int* const s_len = mc.len; int* const s_B = mc.B;             \
    (void)s_base; (void)s_B;                                                                                           \
    Blk b;                                                                                                             \
    b.ncell = min(cfg.C, nocc - grp * cfg.C);                                                                          \
    b.t0 = s_cs[0];                                                                                                    \
    b.t1 = s_cs[b.ncell];                                                                                              \
    stage_nodes_ids<D, WANT_Q, NF>(m, G, cfg, b.ncell, mc, ids, s_rank, SQ, s_X, SU, SA);                               \
    __syncthreads();                                                                                                   \
    const int g_n = grp + gridDim.x, g_nn = g_n + gridDim.x;                                                           \
    int ids_n[NLPS_NID];                                                                                               \
    int4 mt_nn = make_int4(0, 0, 0, 0);                                                                                \
    if (g_n < ngroups) ids_fetch(m, cfg, min(cfg.C, nocc - g_n * cfg.C), mn, ids_n);                                    \
    if (g_nn < ngroups) mt_nn = meta_fetch(G, cfg, nocc, g_nn);
// (the body must end with a __syncthreads(): nobody reads the metadata of this group any more)
#define NLPS_PIPE_END()                                                                                                \
    if (g_nn < ngroups) meta_put(cfg, nocc, P.np, g_nn, mt_nn, mc);                                                     \
    _Pragma("unroll") for (int u = 0; u < NLPS_NID; u++) ids[u] = ids_n[u];                                            \
    { const MetaPtrs t_ = mc; mc = mn; mn = t_; }                                                                      \
  }

__device__ __forceinline__ int cell_of(const int* s_cs, int ncell, int t) {
  int lo = 0, hi = ncell;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (s_cs[mid] <= t) lo = mid; else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------------------
// K0 + K1 (fused): per-particle LME update, explicit predictor, and the cell sums of the lumped mass and
// of the mass-weighted displacement increment.
//   particle phase (thread / particle): tributary__LME__ (LME.c:1019-1099) with the PREVIOUS beta,
//     beta__LME__ (LME.c:177-185), __lambda_Newton_Rapson (LME.c:272-353), __predictor_PARTICLES
//     (U-Verlet.c:229-253).  The unnormalised weights exp(-beta|l|^2 + lambda.l) of the converged
//     evaluation stay in shared memory (CACHE) ...
//   cell phase (warp / cell, lane / 2-ring node): ... so that M_A, sum m_p N_A DU_p (U-Verlet.c:166-225,
//     301-367) over the cell's particles need no second evaluation and no atomics; the partials go to
//     part[(slot of the cell in A's transposed ring, rank(A))], summed per node in a fixed order by k_grid_disp.
// Optional (-DNLPS_LME_LP=2 / -DNLPS_G2P_LP=2, 3D only): the neighbour loops of a particle are split over LP = 2 adjacent
// lanes (mask words w with w % LP == lane % LP), the Newton sums are combined by shuffles, so that an 8-cell group of
// 64 particles fills the 128 threads and the dependent chain per thread halves.  Both lanes of a pair compute
// bit-identical sums (a + b == b + a) and take the same branches.  OFF by default: the 3D kernels are bound by
// shared-memory wavefronts and issue slots, not by the length of the per-thread chain -- measured 5 % (k_lme_p2g) and
// 14 % (k_g2p) SLOWER on the 64^3 cube (gpurun_out/ab_3d2.log, ab_3d3.log).
template <int D, int W>
struct LanesPerParticle { static constexpr int value = (D == 3 && W % 2 == 0) ? 2 : 1; };
template <int LP>
__device__ __forceinline__ double pair_sum(double v, unsigned pm) {
#pragma unroll
  for (int o = 1; o < LP; o <<= 1) v += __shfl_xor_sync(pm, v, o);
  return v;
}
template <int D, int W, bool CACHE>
__global__ void __launch_bounds__(128) k_lme_p2g(MeshDev m, PartDev P, GridDev G, StepParams sp, BlockCfg cfg, int* err,
                                                 int do_predictor, const AlmeDev al) {
  extern __shared__ __align__(16) unsigned char smem[];
  const LayoutA<D, W, CACHE> L(cfg);
  int* s_rank = (int*)(smem + L.rank); unsigned char* s_q = smem + L.q; double* s_X = (double*)(smem + L.X);
  double* s_pa = (double*)(smem + L.pa); double* s_zinv = (double*)(smem + L.zinv);
  double* s_mass = (double*)(smem + L.mass); double* s_ddis = (double*)(smem + L.ddis);
  uint32_t* s_mask = (uint32_t*)(smem + L.mask);
  double* s_px = (double*)(smem + L.px); double* s_plam = (double*)(smem + L.plam); double* s_pbeta = (double*)(smem + L.pbeta);
  double* s_cw = (double*)(smem + L.cw); unsigned char* s_pre = smem + L.pre;
#ifndef NLPS_LME_LP
#define NLPS_LME_LP 1  // measured on the 64^3 cube: 3.54 ms with one lane per particle, 3.72 ms with two
#endif
  constexpr int LP = (CACHE || NLPS_LME_LP < 2) ? 1 : LanesPerParticle<D, W>::value;
  const int NC = (CACHE || LP > 1) ? 0 : cfg.NCA;
  int& s_ovf = *(int*)(smem + L.ovf);  // a particle of the chunk has more neighbours than the compact cache holds: recompute
  double* s_tab = (double*)(smem + L.tab);
  if (threadIdx.x < 32) s_tab[threadIdx.x] = g_exp2tab[threadIdx.x];
  const int nocc = *G.n_occ, ngroups = (nocc + cfg.C - 1) / cfg.C;
  // persistent blocks: each walks over cell groups blockIdx.x, blockIdx.x + gridDim.x, ... (pipelined)
  NLPS_PIPE_BEGIN(true, 0, s_q, nullptr, nullptr)
  const int SL = cfg.SL, np = P.ld;  // np: SoA stride
  for (int tb = b.t0; tb < b.t1; tb += cfg.PCAP) {
    const int nb = min(cfg.PCAP, b.t1 - tb);
    if (threadIdx.x == 0) s_ovf = 0;
    __syncthreads();
    // ---- particle phase
    for (int jt = threadIdx.x; jt < nb * LP; jt += blockDim.x) {
      const int j = jt / LP, sub = jt % LP;
      const unsigned pm = (LP == 1) ? 0u : (((1u << LP) - 1u) << ((threadIdx.x & 31) & ~(LP - 1)));
      (void)pm;
      const int t = tb + j, p = G.plist[t];
      const int ci = cell_of(s_cs, b.ncell, t);
      const int len = s_len[ci];
      const int* rk = s_rank + ci * SL;
      const double* Xc = s_X + (size_t)ci * SL * D;
      double xp[D], lam[D];
#pragma unroll
      for (int i = 0; i < D; i++) { xp[i] = P.x[i * np + p]; lam[i] = P.lam[i * np + p]; }
      const double beta_old = P.beta[p];
      const double mp = P.mass[p];
      // aLME (Nodes/aLME.c): the particle's metric B and cut-off ellipsoid C of this search -- set from h_avg of the closest
      // node by the initialisation (:32-166), convected with the last DF^-1 by every later search (:645-652)
      double Bm[4] = {0.0, 0.0, 0.0, 0.0}, Cm[4] = {0.0, 0.0, 0.0, 0.0};
      const bool alme = (D == 2) && al.bten != nullptr;
      if (D == 2 && alme) {
#pragma unroll
        for (int c = 0; c < 4; c++) { Bm[c] = al.bten[(size_t)c * np + p]; Cm[c] = al.cten[(size_t)c * np + p]; }
        if (!sp.reuse_lists) {
          if (!sp.update_I0) {
            const double hA = m.h_avg[s_B[ci]];
            Bm[0] = Bm[3] = __ddiv_rn(sp.gamma_lme, __dmul_rn(hA, hA));
            Cm[0] = Cm[3] = __ddiv_rn(sp.gamma_lme, __dmul_rn(__dmul_rn(sp.neg_log_tol, hA), hA));
            Bm[1] = Bm[2] = Cm[1] = Cm[2] = 0.0;
          } else {
            double DFm[4], Fi[4];
#pragma unroll
            for (int c = 0; c < 4; c++) DFm[c] = P.DF[(size_t)c * np + p];
            if (inverse<2>(DFm, Fi) == 0.0) { if (sub == 0) latch_error(err, NLPS_ERR_SINGULAR_DF, P.orig[p]); }
            alme_push_forward(Fi, Bm);
            alme_push_forward(Fi, Cm);
          }
          if (sub == 0) {
#pragma unroll
            for (int c = 0; c < 4; c++) { al.bten[(size_t)c * np + p] = Bm[c]; al.cten[(size_t)c * np + p] = Cm[c]; }
          }
        }
      }
      const MetricB MB = {Bm[0], Bm[1] + Bm[2], Bm[3], alme ? 1 : 0};
      double pv[D], pa_[D];  // requested now, consumed by the predictor after the Newton loop
#pragma unroll
      for (int i = 0; i < D; i++) {
        pv[i] = do_predictor ? P.vel[i * np + p] : (sp.proj ? sp.proj[i * np + p] : P.ddis[i * np + p]);
        pa_[i] = do_predictor ? P.acc[i * np + p] : 0.0;
      }
      const double Ra = __dsqrt_rn(__ddiv_rn(sp.neg_log_tol, beta_old));  // LME.c:1052
      const double sstar = sstar_from_Ra(Ra);
      uint32_t mk[W];
#pragma unroll
      for (int w = 0; w < W; w++) mk[w] = 0u;
      int n = 0;
      if (sp.reuse_lists) {
#pragma unroll
        for (int w = 0; w < W; w++) {
          mk[w] = P.mask[(size_t)w * np + p];
          n += __popc(mk[w]);
          if (LP > 1 && (w % LP) != sub) mk[w] = 0u;  // this lane walks its own words only
        }
      } else {
        if (LP == 1) {
          for (int k = 0; k < len; k++) {
            if (rk[k] < 0) continue;  // inactive node
            double l[D];
            const double s = dist2_exact<D>(xp, Xc + k * D, l);
            const bool in = (D == 2 && alme) ? alme_distance(Cm, l) <= 1.0 : s <= sstar;  // tributary__aLME__ (aLME.c:811-872)
            if (in) { mk[k >> 5] |= 1u << (k & 31); n++; }
          }
        } else {
#pragma unroll
          for (int w = 0; w < W; w++) {
            if ((w % LP) != sub) continue;
            const int k1 = min(len, 32 * w + 32);
            for (int k = 32 * w; k < k1; k++) {
              if (rk[k] < 0) continue;
              double l[D];
              const double s = dist2_exact<D>(xp, Xc + k * D, l);
              const bool in = (D == 2 && alme) ? alme_distance(Cm, l) <= 1.0 : s <= sstar;
              if (in) { mk[w] |= 1u << (k & 31); n++; }
            }
          }
#pragma unroll
          for (int o = 1; o < LP; o <<= 1) n += __shfl_xor_sync(pm, n, o);
        }
#pragma unroll
        for (int w = 0; w < W; w++)
          if (LP == 1 || (w % LP) == sub) P.mask[(size_t)w * np + p] = mk[w];
        if (sub == 0) P.nnodes[p] = n;
      }
      if (CACHE && !DenseSlots<D, W>::value) {  // slots that are not neighbours carry weight 0: the cell phase needs no mask test
#pragma unroll
        for (int w = 0; w < W; w++) {
          const int rem = len - 32 * w;
          uint32_t nm = ~mk[w] & (rem >= 32 ? 0xffffffffu : (rem > 0 ? (1u << rem) - 1u : 0u));
          while (nm) { s_pa[(size_t)j * SL + w * 32 + __ffs(nm) - 1] = 0.0; nm &= nm - 1; }
        }
      }
      bool ok = true;
      if (n < D + 1) { if (sub == 0) latch_error(err, NLPS_ERR_FEW_NEIGHBOURS, P.orig[p]); ok = false; }
      const double h = m.h_avg[s_B[ci]];
      const double beta = sp.reuse_lists ? beta_old : __ddiv_rn(sp.gamma_lme, __dmul_rn(h, h));
      if (!sp.reuse_lists && sub == 0) P.beta[p] = beta;
      // Newton on lambda
      int NumIter = 0;
      double Z = 1.0;
      while (ok && NumIter <= sp.max_iter_lme) {
        double r[D], JJ[D * D];
        Z = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) r[i] = 0.0;
#pragma unroll
        for (int i = 0; i < D * D; i++) JJ[i] = 0.0;
        int ord = 0;  // ordinal of the neighbour in ascending slot order (compact cache)
        const bool keep = LP == 1 && !CACHE && NC > 0 && n <= NC;
        for_slots<D, W>(mk, len, [&](int k0, int k1, double w0, double w1) {
          double l0[D], l1[D], X0[D], X1[D], ll0 = 0.0, lx0 = 0.0, ll1 = 0.0, lx1 = 0.0;
          ldsvec<D>(Xc + k0 * D, X0);
          ldsvec<D>(Xc + k1 * D, X1);
#pragma unroll
          for (int i = 0; i < D; i++) {
            l0[i] = xp[i] - X0[i];
            l1[i] = xp[i] - X1[i];
            ll0 += l0[i] * l0[i];
            ll1 += l1[i] * l1[i];
            lx0 += l0[i] * lam[i];
            lx1 += l1[i] * lam[i];
          }
          const double e0 = fexp(-metric_q<D>(MB, beta, ll0, l0) + lx0, s_tab) * w0;
          const double e1 = fexp(-metric_q<D>(MB, beta, ll1, l1) + lx1, s_tab) * w1;
          if (CACHE) {
            s_pa[(size_t)j * SL + k0] = e0;
            if (k1 != k0) s_pa[(size_t)j * SL + k1] = e1;
          }
          if (keep) {
            s_cw[(size_t)j * NC + ord] = e0;
            if (k1 != k0) s_cw[(size_t)j * NC + ord + 1] = e1;
            ord += (k1 != k0) ? 2 : 1;  // a mask word with an odd number of bits ends on a single
          }
          Z += e0;
#pragma unroll
          for (int i = 0; i < D; i++) {
            r[i] += e0 * l0[i];
#pragma unroll
            for (int jj = i; jj < D; jj++) JJ[i * D + jj] += e0 * l0[i] * l0[jj];
          }
          Z += e1;
#pragma unroll
          for (int i = 0; i < D; i++) {
            r[i] += e1 * l1[i];
#pragma unroll
            for (int jj = i; jj < D; jj++) JJ[i * D + jj] += e1 * l1[i] * l1[jj];
          }
        });
        if (LP > 1) {  // the two halves of the neighbour sums
          Z = pair_sum<LP>(Z, pm);
#pragma unroll
          for (int i = 0; i < D; i++) {
            r[i] = pair_sum<LP>(r[i], pm);
#pragma unroll
            for (int jj = i; jj < D; jj++) JJ[i * D + jj] = pair_sum<LP>(JJ[i * D + jj], pm);
          }
        }
        const double Zi = 1.0 / Z;
        double nr = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) { r[i] *= Zi; nr += r[i] * r[i]; }
        nr = sqrt(nr);
        if (nr > sp.tol_wrapper) {
#pragma unroll
          for (int i = 0; i < D; i++)
#pragma unroll
            for (int jj = i; jj < D; jj++) {
              JJ[i * D + jj] = JJ[i * D + jj] * Zi - r[i] * r[jj];
              JJ[jj * D + i] = JJ[i * D + jj];
            }
          if (rcond_as_reference<D>(JJ) < 1E-8) { ok = false; if (sub == 0) latch_error(err, NLPS_ERR_SINGULAR_HESSIAN, P.orig[p]); break; }
          double Ji[D * D];
          inverse<D>(JJ, Ji);
#pragma unroll
          for (int i = 0; i < D; i++) {
            double dl = 0.0;
#pragma unroll
            for (int jj = 0; jj < D; jj++) dl += Ji[i * D + jj] * r[jj];
            lam[i] -= dl;
          }
          NumIter++;
        } else {
          break;
        }
      }
      if (ok && NumIter >= sp.max_iter_lme && sub == 0) latch_error(err, NLPS_ERR_NEWTON_LME, P.orig[p]);
#pragma unroll
      for (int w = 0; w < W; w++)
        if (LP == 1 || (w % LP) == sub) s_mask[j * W + w] = ok ? mk[w] : 0u;
      if (sub != 0) continue;  // the first lane of the pair writes the particle's results
#pragma unroll
      for (int i = 0; i < D; i++) P.lam[i * np + p] = lam[i];
      // predictor (gamma = 0.5, U-Verlet.c:76,248)
#pragma unroll
      for (int i = 0; i < D; i++) {
        double dd;
        if (do_predictor) {
          const double v = pv[i], a = pa_[i];
          dd = sp.dt * v + 0.5 * (sp.dt * sp.dt) * a;
          P.ddis[i * np + p] = dd;
          P.vel[i * np + p] = v + (1 - 0.5) * sp.dt * a;
        } else {
          dd = pv[i];
        }
        s_ddis[j * D + i] = dd;
        if (!CACHE) { s_px[j * D + i] = xp[i]; s_plam[j * D + i] = lam[i]; }
      }
      if (!CACHE) s_pbeta[j] = beta;
      if (!CACHE && NC > 0) {
        if (n > NC) s_ovf = 1;
        int acc = 0;
#pragma unroll
        for (int w = 0; w < W; w++) { s_pre[j * W + w] = (unsigned char)acc; acc += __popc(mk[w]); }
      }
      s_zinv[j] = ok ? 1.0 / Z : 0.0;
      s_mass[j] = mp;
      if (CACHE) {
        // the cell phase sums weight * (m/Z) and weight * (m/Z) * DU_p
        const double wgt = ok ? mp / Z : 0.0;
        s_zinv[j] = wgt;
#pragma unroll
        for (int i = 0; i < D; i++) s_ddis[j * D + i] *= wgt;
        if (!ok)
          for (int k = 0; k < len; k++) s_pa[(size_t)j * SL + k] = 0.0;
      }
    }
    __syncthreads();
    // ---- cell phase: one thread per (cell, 2-ring node) pair sums over the cell's particles
    for (int q = threadIdx.x; q < b.ncell * SL; q += blockDim.x) {
      // weights cached (2D): slot fastest.  Recomputed weights (3D): cell fastest, so that the lanes of a warp ask for the
      // same slot of different cells -- on regular clouds the same neighbour pattern, i.e. no divergence at the mask test
      int c, k;
      if (CACHE || !(cfg.cellfast & 1)) { c = (int)(((unsigned)q * cfg.magic) >> 21); k = q - c * SL; }
      else { k = q / b.ncell; c = q - k * b.ncell; }
      const int e = c * SL + k;
      const int rank = s_rank[e];
      if (rank < 0) continue;
      const int ja = max(s_cs[c], tb) - tb, jb = min(s_cs[c + 1], tb + nb) - tb;
      if (ja >= jb) continue;
      const bool first = s_cs[c] >= tb;
      const int kw = k >> 5;
      const uint32_t kb = 1u << (k & 31);
      double a0 = 0.0, a[D];
#pragma unroll
      for (int i = 0; i < D; i++) a[i] = 0.0;
      if (CACHE) {
        for (int j = ja; j < jb; j++) {
          const double pa = s_pa[(size_t)j * SL + k];
          a0 += pa * s_zinv[j];
#pragma unroll
          for (int i = 0; i < D; i++) a[i] += pa * s_ddis[j * D + i];
        }
      } else {
        const bool use_cw = NC > 0 && !s_ovf;
        for (int j = ja; j < jb; j++) {
          const uint32_t mw = s_mask[j * W + kw];
          if (!(mw & kb)) continue;
          double wexp;
          if (use_cw) {
            wexp = s_cw[(size_t)j * NC + s_pre[j * W + kw] + __popc(mw & (kb - 1u))];
          } else {
            double ll = 0.0, lx = 0.0;
#pragma unroll
            for (int i = 0; i < D; i++) {
              const double l = s_px[j * D + i] - s_X[(size_t)e * D + i];
              ll += l * l;
              lx += l * s_plam[j * D + i];
            }
            wexp = fexp(-s_pbeta[j] * ll + lx, s_tab);
          }
          const double mN = wexp * s_zinv[j] * s_mass[j];
          a0 += mN;
#pragma unroll
          for (int i = 0; i < D; i++) a[i] += mN * s_ddis[j * D + i];
        }
      }
      // (cell-major: slot-fastest pairs write consecutive 32-byte records of the cell's run; GridDev::part)
      double* dst = G.cm_sl ? G.part + ((size_t)(grp * cfg.C + c) * G.cm_sl + s_q[e]) * 4
                            : G.part + ((size_t)s_q[e] * G.max_act + rank) * (1 + D);
      if (first) {
        dst[0] = a0;
#pragma unroll
        for (int i = 0; i < D; i++) dst[1 + i] = a[i];
      } else {
        dst[0] += a0;
#pragma unroll
        for (int i = 0; i < D; i++) dst[1 + i] += a[i];
      }
    }
    __syncthreads();
  }
  NLPS_PIPE_END()  // cell groups
}

// Stage 2 + G1 (node kernel): M_A = sum_p N_A m_p (U-Verlet.c:166-225); DU_A = sum_p m_p N_A DU_p / M_A
// (U-Verlet.c:301-367) as a fixed-order sum of the cell partials (coalesced: slot-major layout);
// Dirichlet overwrite (U-Verlet.c:458-526) and restricted-DOF flags (Nodes-Tools.c:70-156).
struct BcDev {
  const int *node_ptr, *node_bnd;  // CSR: node -> boundary ids in boundary order
  const int* bnd_dim;
  const int* dir;      // [b][k][step] flattened with stride maxdim*nsteps
  const double* val;
  int maxdim, nsteps, nb;
};

__global__ void k_build_r2ts(int nn, const int* r2p, const int* r2i, const unsigned char* r2q, const int* r2tp, unsigned char* r2ts,
                             unsigned char* r2pos, int sorted) {
  const int B = blockIdx.x * blockDim.x + threadIdx.x;
  if (B >= nn) return;
  const int b0 = r2p[B], len = r2p[B + 1] - b0;
  for (int s_ = 0; s_ < len; s_++) {
    const int A = r2i[b0 + s_];
    int pos = s_;
    if (sorted) {  // rank of A among the node ids of the ring (ids are distinct)
      pos = 0;
      for (int k = 0; k < len; k++) pos += r2i[b0 + k] < A;
    }
    r2ts[r2tp[A] + r2q[b0 + s_]] = (unsigned char)pos;
    r2pos[b0 + s_] = (unsigned char)pos;
  }
}
__global__ void k_set_flags(const int* ids, int n, unsigned char* flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[ids[i]] = 1;
}
// MODE 0: everything (single slab).  MODE 1: sums only, M and MOM stored (the slab halo exchange adds the
// neighbour slabs' sums on the shared nodes).  MODE 2: division + Dirichlet from the stored sums -- over all active
// nodes, or (band lists given) over the nodes of the halo bands only.  MODE 3 (slabs): MODE 0 for the nodes outside
// the halo bands, MODE 1 for the band nodes, which MODE 2 finishes after the exchange: the second pass touches a few
// thousand nodes instead of all of them.
// Cell-major partial sums (G.cm_sl > 0, 3D): NV doubles of every non-zero record (cell, slot of node A) over the occupied
// cells of A's transposed 2-ring, a WARP per node: the lanes take the ring positions q, q + 32, ..., every lane has its
// (scattered, 32-byte) record loads in flight at once, and the lane sums meet in a fixed butterfly (deterministic).
template <int NV, int LPN>
__device__ __forceinline__ void node_sums_cm(const MeshDev& m, const GridDev& G, int A, int t, int lane, double* out) {
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  const int q0 = m.r2tp[A], nq = m.r2tp[A + 1] - q0;
  // the dependent chain ring position -> cell -> (rank of the cell, non-zero mask) -> record is walked level by level
  // for four ring positions of the lane at once: the loads of a level are independent and in flight together
  constexpr int R = 4;
  for (int qb = 0; qb < nq; qb += LPN * R) {
    int B[R], sl[R], u[R];
    bool on[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int q = qb + LPN * r + lane;
      on[r] = q < nq && ((G.occm[(size_t)(q >> 5) * G.max_act + t] >> (q & 31)) & 1u);
      B[r] = on[r] ? m.r2ti[q0 + q] : 0;
      sl[r] = on[r] ? (int)m.r2ts[q0 + q] : 0;
    }
#pragma unroll
    for (int r = 0; r < R; r++) u[r] = on[r] ? G.occ_pos[B[r]] : 0;
#pragma unroll
    for (int r = 0; r < R; r++)  // a zero record (the slot is no particle's neighbour) is skipped
      on[r] = on[r] && (!G.cum || ((G.cum[(size_t)u[r] * G.cm_w + (sl[r] >> 5)] >> (sl[r] & 31)) & 1u));
    double2 v01[R], v23[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const double* src = G.part + ((size_t)u[r] * G.cm_sl + sl[r]) * 4;
      v01[r] = on[r] ? *reinterpret_cast<const double2*>(src) : make_double2(0.0, 0.0);
      v23[r] = on[r] ? *reinterpret_cast<const double2*>(src + 2) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int r = 0; r < R; r++) { a[0] += v01[r].x; a[1] += v01[r].y; a[2] += v23[r].x; a[3] += v23[r].y; }
  }
#pragma unroll
  for (int o = LPN / 2; o > 0; o >>= 1)  // (the lanes of a node are LPN consecutive lanes of the warp)
#pragma unroll
    for (int i = 0; i < NV; i++) a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
#pragma unroll
  for (int i = 0; i < NV; i++) out[i] = a[i];
}

template <int D, int MODE, int LPN = 1>  // LPN: lanes per node (cell-major partial sums: 32 in 3D, 8 in 2D; 1 = slot-major)
__global__ void __launch_bounds__(128) k_grid_disp(MeshDev m, GridDev G, BcDev bc, int step, const int* ids0 = nullptr,
                                                   int n0 = 0, const int* ids1 = nullptr, int n1 = 0) {
  int t = (blockIdx.x * blockDim.x + threadIdx.x) / LPN;
  const int lane = threadIdx.x % LPN;
  int A;
  bool live = true;
  if (MODE == 2 && (ids0 || ids1)) {
    if (t >= n0 + n1) return;
    A = t < n0 ? ids0[t] : ids1[t - n0];
    if (!G.active[A]) return;
  } else {
    const int na = *G.n_active;
    if (LPN == 1 ? t >= na : na == 0) return;
    // several nodes per warp (LPN < 32): the lanes of a node past the end keep running on node 0 (the shuffles below
    // are warp-wide) and drop out after the sums
    live = t < na;
    if (!live) t = 0;
    A = G.act_list[t];
  }
  double mom[D], M = 0.0;
#pragma unroll
  for (int i = 0; i < D; i++) mom[i] = 0.0;
  if (MODE == 2) {
    M = G.M[A];
#pragma unroll
    for (int i = 0; i < D; i++) mom[i] = G.MOM[(size_t)A * D + i];
  }
  if (LPN > 1) {
    if (MODE != 2) {
      double sums[4];
      node_sums_cm<1 + D, LPN>(m, G, A, t, lane, sums);
      M = sums[0];
#pragma unroll
      for (int i = 0; i < D; i++) mom[i] = sums[1 + i];
    }
    if (lane != 0 || !live) return;
  }
  for (int w = 0; LPN == 1 && MODE != 2 && w < G.w2t; w++) {
    uint32_t mm = G.occm[(size_t)w * G.max_act + t];
    while (mm) {
      const int q = w * 32 + __ffs(mm) - 1;
      mm &= mm - 1;
      const double* src = G.part + ((size_t)q * G.max_act + t) * (1 + D);
      M += src[0];
#pragma unroll
      for (int i = 0; i < D; i++) mom[i] += src[1 + i];
    }
  }
  if (MODE == 1 || (MODE == 3 && G.band[A])) {
    G.M[A] = M;
#pragma unroll
    for (int i = 0; i < D; i++) G.MOM[(size_t)A * D + i] = mom[i];
    return;
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
// K2 + K3 (fused): kinematics, stress, and the cell sums of the nodal forces.
//   particle phase: DF = I + sum_A DU_A (x) gradN_A with gradN_a = -p_a J^-1 l_a (compute-Strains.c:20-44,
//     LME.c:836-891); F_n1 = DF F_n (compute-Strains.c:76-105); J > 0 (U-Verlet.c:608-613);
//     rho /= det DF (U-Verlet.c:630-632); stress (Constitutive.c:18-258);
//     G_p = V0 tau DF^-T J^-1, so that f_A = sum_p N_A (G_p l_A + t_p) == -V0 tau (DF^-T gradN_A) + N_A T A0
//     (U-Newmark-beta.c:1257-1374 with Shape-Functions.c:405-448; tractions U-Verlet.c:805-902).
//   cell phase: that sum over the cell's particles, weights from shared memory.
// MAT: compile-time material law when every particle uses the same one (keeps the Matsuoka-Nakai
// Newton out of the register budget of the other laws); -1 = mixed, dispatched per particle.
template <int D, int W, int MAT, bool CACHE>
__global__ void __launch_bounds__(128, (MAT == 0 || MAT == 1) ? 3 : 2) k_kin_force(MeshDev m, PartDev P, GridDev G,
                                                                                StepParams sp, BlockCfg cfg, int* err,
                                                                                int has_traction,
                                                                                const __grid_constant__ MatTable mt,
                                                                                const AlmeDev al) {
  extern __shared__ __align__(16) unsigned char smem[];
  const LayoutB<D, W, CACHE> L(cfg);
  int* s_rank = (int*)(smem + L.rank); unsigned char* s_q = smem + L.q; double* s_X = (double*)(smem + L.X);
  double* s_U = (double*)(smem + L.U); double* s_pa = (double*)(smem + L.pa); double* s_zinv = (double*)(smem + L.zinv);
  double* s_G = (double*)(smem + L.G); double* s_px = (double*)(smem + L.px); double* s_trac = (double*)(smem + L.trac);
  uint32_t* s_mask = (uint32_t*)(smem + L.mask);
  double* s_plam = (double*)(smem + L.plam); double* s_pbeta = (double*)(smem + L.pbeta);
  double* s_tab = (double*)(smem + L.tab);
  if (threadIdx.x < 32) s_tab[threadIdx.x] = g_exp2tab[threadIdx.x];
  const int nocc = *G.n_occ, ngroups = (nocc + cfg.C - 1) / cfg.C;
  // persistent blocks: each walks over cell groups blockIdx.x, blockIdx.x + gridDim.x, ... (pipelined)
  NLPS_PIPE_BEGIN(true, 1, s_q, s_U, nullptr)
  const int SL = cfg.SL, np = P.ld;  // np: SoA stride
  constexpr int T = (D == 2) ? 5 : 9;
  for (int tb = b.t0; tb < b.t1; tb += cfg.PCAP) {
    const int nb = min(cfg.PCAP, b.t1 - tb);
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
      const int t = tb + j, p = G.plist[t];
      const int ci = cell_of(s_cs, b.ncell, t);
      const double* Xc = s_X + (size_t)ci * SL * D;
      const double* Uc = s_U + (size_t)ci * SL * D;
      double xp[D], lam[D];
#pragma unroll
      for (int i = 0; i < D; i++) { xp[i] = P.x[i * np + p]; lam[i] = P.lam[i * np + p]; }
      const double beta = P.beta[p];
      const MetricB MB = metric_load<D>(al, P.ld, p);  // aLME: the particle's metric instead of beta
      uint32_t mk[W];
#pragma unroll
      for (int w = 0; w < W; w++) mk[w] = P.mask[(size_t)w * np + p];
      if (CACHE && !DenseSlots<D, W>::value) {  // slots that are not neighbours carry weight 0: the cell phase needs no mask test
        const int len = s_len[ci];
#pragma unroll
        for (int w = 0; w < W; w++) {
          const int rem = len - 32 * w;
          uint32_t nm = ~mk[w] & (rem >= 32 ? 0xffffffffu : (rem > 0 ? (1u << rem) - 1u : 0u));
          while (nm) { s_pa[(size_t)j * SL + w * 32 + __ffs(nm) - 1] = 0.0; nm &= nm - 1; }
        }
      }
      // particle state requested now, consumed after the neighbour loop (loads overlap the loop)
      double Fn[D * D], be[T];
#pragma unroll
      for (int i = 0; i < D * D; i++) Fn[i] = P.F_n[(size_t)i * np + p];
      const double rho_p = P.rho[p], V0 = P.vol0[p];
      const int mid = P.matidx[p];
      const bool plastic = (MAT >= 0) ? (MAT != NLPS_MAT_NEO_HOOKEAN_WRIGGERS) : true;
      double eps = 0.0, kap = 0.0, back[3] = {0.0, 0.0, 0.0};
      if (MAT < 0 && P.back) {
#pragma unroll
        for (int i = 0; i < 3; i++) back[i] = P.back[(size_t)i * np + p];
      }
      if (plastic) {
#pragma unroll
        for (int i = 0; i < T; i++) be[i] = P.be_n[(size_t)i * np + p];
        eps = P.eps_n[p];
        kap = P.kap_n[p];
      }
      double Z = 0.0, r[D], JJ[D * D], Bm[D * D];
#pragma unroll
      for (int i = 0; i < D; i++) r[i] = 0.0;
#pragma unroll
      for (int i = 0; i < D * D; i++) { JJ[i] = 0.0; Bm[i] = 0.0; }
      for_slots<D, W>(mk, s_len[ci], [&](int k0, int k1, double w0, double w1) {
        double l0[D], l1[D], X0[D], X1[D], U0[D], U1[D], ll0 = 0.0, lx0 = 0.0, ll1 = 0.0, lx1 = 0.0;
        ldsvec<D>(Xc + k0 * D, X0);
        ldsvec<D>(Xc + k1 * D, X1);
        ldsvec<D>(Uc + k0 * D, U0);
        ldsvec<D>(Uc + k1 * D, U1);
#pragma unroll
        for (int i = 0; i < D; i++) {
          l0[i] = xp[i] - X0[i];
          l1[i] = xp[i] - X1[i];
          ll0 += l0[i] * l0[i];
          ll1 += l1[i] * l1[i];
          lx0 += l0[i] * lam[i];
          lx1 += l1[i] * lam[i];
        }
        const double e0 = fexp(-metric_q<D>(MB, beta, ll0, l0) + lx0, s_tab) * w0;
        const double e1 = fexp(-metric_q<D>(MB, beta, ll1, l1) + lx1, s_tab) * w1;
        if (CACHE) {
          s_pa[(size_t)j * SL + k0] = e0;
          if (k1 != k0) s_pa[(size_t)j * SL + k1] = e1;
        }
        Z += e0;
#pragma unroll
        for (int i = 0; i < D; i++) {
          r[i] += e0 * l0[i];
          const double eu = e0 * U0[i];
#pragma unroll
          for (int jj = 0; jj < D; jj++) {
            if (jj >= i) JJ[i * D + jj] += e0 * l0[i] * l0[jj];
            Bm[i * D + jj] += eu * l0[jj];
          }
        }
        Z += e1;
#pragma unroll
        for (int i = 0; i < D; i++) {
          r[i] += e1 * l1[i];
          const double eu = e1 * U1[i];
#pragma unroll
          for (int jj = 0; jj < D; jj++) {
            if (jj >= i) JJ[i * D + jj] += e1 * l1[i] * l1[jj];
            Bm[i * D + jj] += eu * l1[jj];
          }
        }
      });
      const double Zi = 1.0 / Z;
#pragma unroll
      for (int i = 0; i < D; i++) r[i] *= Zi;
#pragma unroll
      for (int i = 0; i < D; i++)
#pragma unroll
        for (int jj = i; jj < D; jj++) {
          JJ[i * D + jj] = JJ[i * D + jj] * Zi - r[i] * r[jj];
          JJ[jj * D + i] = JJ[i * D + jj];
        }
      double Ji[D * D];
      inverse<D>(JJ, Ji);
      double DF[D * D], Fn1[D * D];
#pragma unroll
      for (int i = 0; i < D; i++)
#pragma unroll
        for (int jj = 0; jj < D; jj++) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < D; k++) s += Bm[i * D + k] * Ji[k * D + jj];
          DF[i * D + jj] = ((i == jj) ? 1.0 : 0.0) - s * Zi;
        }
#pragma unroll
      for (int i = 0; i < D; i++)
#pragma unroll
        for (int jj = 0; jj < D; jj++) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < D; k++) s += DF[i * D + k] * Fn[k * D + jj];
          Fn1[i * D + jj] = s;
          P.F_n1[(size_t)(i * D + jj) * np + p] = s;
          P.DF[(size_t)(i * D + jj) * np + p] = DF[i * D + jj];
        }
      const double J1 = det<D>(Fn1);
      P.J_n1[p] = J1;
      bool ok = true;
      double Gp[D * D];
#pragma unroll
      for (int i = 0; i < D * D; i++) Gp[i] = 0.0;
      if (J1 <= 0.0) { latch_error(err, NLPS_ERR_NEGATIVE_JACOBIAN, P.orig[p]); ok = false; }
      if (ok) {
        const double dJ = det<D>(DF);
        if (!sp.implicit) P.rho[p] = rho_p / dJ;  // the implicit scheme updates rho once, after convergence
        // constitutive update
        const MatParams& mat = mt.m[mid];
        double tau[T], Wp = 0.0;
        const int mtype = (MAT >= 0) ? MAT : mat.type;
        int st = 0;
        if (mtype == NLPS_MAT_NEO_HOOKEAN_WRIGGERS) {
          stress_neo_hookean<D>(mat, Fn1, J1, tau, Wp);
        } else {
          double cep[D * D];
          if (MAT == NLPS_MAT_DRUCKER_PRAGER) st = stress_drucker_prager<D>(mat, sp.rp, DF, be, eps, kap, tau, Wp, cep);
          else if (MAT == NLPS_MAT_MATSUOKA_NAKAI) st = stress_matsuoka_nakai<D>(mat, sp.rp, DF, be, eps, kap, tau, Wp, cep);
          else st = stress_with_history<D>(mtype, mat, sp.rp, DF, Fn1, be, eps, kap, back, tau, Wp, cep);
          if (st != 0) { latch_error(err, st, P.orig[p]); ok = false; }
          if (ok && (MAT >= 0 || mat_has_history(mtype))) {
#pragma unroll
            for (int i = 0; i < T; i++) P.be_n1[(size_t)i * np + p] = be[i];
            P.eps_n1[p] = eps;
            if (MAT >= 0 || mtype != NLPS_MAT_VON_MISES) P.kap_n1[p] = kap;  // Von-Mises never touches Kappa
            if (MAT < 0 && mtype == NLPS_MAT_VON_MISES && P.back) {
#pragma unroll
              for (int i = 0; i < 3; i++) P.back[(size_t)i * np + p] = back[i];
            }
            if (sp.rp.want_cep)
#pragma unroll
              for (int i = 0; i < D * D; i++) P.cep[(size_t)i * np + p] = cep[i];
          }
        }
        if (ok) {
#pragma unroll
          for (int i = 0; i < T; i++) P.stress[(size_t)i * np + p] = tau[i];
          P.W[p] = Wp;
          // force operator G = V0 * tau * DF^-T * J^-1
          double DFi[D * D];
          const double dd = inverse<D>(DF, DFi);
          if (dd == 0.0) { latch_error(err, NLPS_ERR_SINGULAR_DF, P.orig[p]); ok = false; }
          double tA[D * D];
#pragma unroll
          for (int i = 0; i < D; i++)
#pragma unroll
            for (int jj = 0; jj < D; jj++) {
              double s = 0.0;
#pragma unroll
              for (int k = 0; k < D; k++) s += tau[i * D + k] * DFi[jj * D + k];  // tau * DF^-T
              tA[i * D + jj] = s;
            }
#pragma unroll
          for (int i = 0; i < D; i++)
#pragma unroll
            for (int jj = 0; jj < D; jj++) {
              double s = 0.0;
#pragma unroll
              for (int k = 0; k < D; k++) s += tA[i * D + k] * Ji[k * D + jj];
              Gp[i * D + jj] = V0 * s;
            }
        }
      }
#pragma unroll
      for (int i = 0; i < D * D; i++) s_G[j * D * D + i] = Gp[i];
#pragma unroll
      for (int i = 0; i < D; i++) {
        s_px[j * D + i] = xp[i];
        s_trac[j * D + i] = has_traction ? P.trac[(size_t)i * np + p] : 0.0;
        if (!CACHE) s_plam[j * D + i] = lam[i];
      }
      if (!CACHE) s_pbeta[j] = beta;
      s_zinv[j] = ok ? Zi : 0.0;
#pragma unroll
      for (int w = 0; w < W; w++) s_mask[j * W + w] = ok ? mk[w] : 0u;
      if (CACHE && !ok)
        for (int k = 0; k < s_len[ci]; k++) s_pa[(size_t)j * SL + k] = 0.0;
    }
    __syncthreads();
    // ---- cell phase: f_A partials, one thread per (cell, 2-ring node) pair
    for (int q = threadIdx.x; q < b.ncell * SL; q += blockDim.x) {
      // weights cached (2D): slot fastest.  Recomputed weights (3D): cell fastest, so that the lanes of a warp ask for the
      // same slot of different cells -- on regular clouds the same neighbour pattern, i.e. no divergence at the mask test
      int c, k;
      if (CACHE || !(cfg.cellfast & 2)) { c = (int)(((unsigned)q * cfg.magic) >> 21); k = q - c * SL; }
      else { k = q / b.ncell; c = q - k * b.ncell; }
      const int e = c * SL + k;
      const int rank = s_rank[e];
      if (rank < 0) continue;
      const int ja = max(s_cs[c], tb) - tb, jb = min(s_cs[c + 1], tb + nb) - tb;
      if (ja >= jb) continue;
      const bool first = s_cs[c] >= tb;
      const int kw = k >> 5;
      const uint32_t kb = 1u << (k & 31);
      double XA[D], f[D];
#pragma unroll
      for (int i = 0; i < D; i++) { XA[i] = s_X[(size_t)e * D + i]; f[i] = 0.0; }
      for (int j = ja; j < jb; j++) {
        if (!CACHE && !(s_mask[j * W + kw] & kb)) continue;
        double l[D], N;
#pragma unroll
        for (int i = 0; i < D; i++) l[i] = s_px[j * D + i] - XA[i];
        if (CACHE) {
          N = s_pa[(size_t)j * SL + k] * s_zinv[j];
        } else {
          double ll = 0.0, lx = 0.0;
#pragma unroll
          for (int i = 0; i < D; i++) { ll += l[i] * l[i]; lx += l[i] * s_plam[j * D + i]; }
          N = fexp(-s_pbeta[j] * ll + lx, s_tab) * s_zinv[j];
        }
#pragma unroll
        for (int i = 0; i < D; i++) {
          double gl = s_trac[j * D + i];
#pragma unroll
          for (int kk = 0; kk < D; kk++) gl += s_G[j * D * D + i * D + kk] * l[kk];
          f[i] += N * gl;
        }
      }
      double* dst = G.cm_sl ? G.part + ((size_t)(grp * cfg.C + c) * G.cm_sl + s_q[e]) * 4
                            : G.part + ((size_t)s_q[e] * G.max_act + rank) * D;
      if (first) {
#pragma unroll
        for (int i = 0; i < D; i++) dst[i] = f[i];
      } else {
#pragma unroll
        for (int i = 0; i < D; i++) dst[i] += f[i];
      }
    }
    __syncthreads();
  }
  NLPS_PIPE_END()  // cell groups
}

// Neumann tractions: per loaded particle t_p = sum_loads T(step) * A0_p, A0 = Vol_0 / thickness
// in 2D, Phi.Area_0 in 3D (U-Verlet.c:826-869), into P.trac (zero for unloaded particles).
struct NeuDev {
  int n_entries;        // flattened (load, particle) pairs, load-major; particle = the caller's id
  const int* part;
  const int* load;
  const int* load_dim;
  const int* dir;
  const double* val;
  int maxdim, nsteps;
};
template <int D>
__global__ void k_traction(PartDev P, NeuDev nu, double thickness, int step) {
  // entries of different loads may hit the same particle: accumulate with fp64 global atomics (tiny set)
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nu.n_entries) return;
  int p = P.inv[nu.part[e]], b = nu.load[e];
  if (p < 0) return;  // the particle lives in another slab
  // 2D: Vol_0 / Thickness_Plain_Stress; 3D: Phi.Area_0 (U-Verlet.c:844-849)
  const double A0 = (D == 3) ? P.area0[p] : P.vol0[p] / thickness;
  for (int k = 0; k < nu.load_dim[b] && k < D; k++) {
    size_t o = ((size_t)b * nu.maxdim + k) * nu.nsteps + step;
    if (nu.dir[o] == 1) atomicAdd(&P.trac[(size_t)k * P.ld + p], nu.val[o] * A0);
  }
}

// K3 stage 2 + G2 (node kernel): f_A = sum of cell partials; a_A = g + f_A / M_A on free DOFs, 0 on
// restricted ones (U-Verlet.c:947-958; gravity as U-Newmark-beta.c:1539-1543).
template <int D, int MODE, int LPN = 1>  // MODE, LPN as in k_grid_disp
__global__ void __launch_bounds__(128) k_grid_acc(MeshDev m, GridDev G, const double* grav, int nsteps, int step,
                                                  const int* ids0 = nullptr, int n0 = 0, const int* ids1 = nullptr, int n1 = 0) {
  int t = (blockIdx.x * blockDim.x + threadIdx.x) / LPN;
  const int lane = threadIdx.x % LPN;
  int A;
  bool live = true;
  if (MODE == 2 && (ids0 || ids1)) {
    if (t >= n0 + n1) return;
    A = t < n0 ? ids0[t] : ids1[t - n0];
    if (!G.active[A]) return;
  } else {
    const int na = *G.n_active;
    if (LPN == 1 ? t >= na : na == 0) return;
    // several nodes per warp (LPN < 32): the lanes of a node past the end keep running on node 0 (the shuffles below
    // are warp-wide) and drop out after the sums
    live = t < na;
    if (!live) t = 0;
    A = G.act_list[t];
  }
  double f[D];
#pragma unroll
  for (int i = 0; i < D; i++) f[i] = (MODE == 2) ? G.F[(size_t)A * D + i] : 0.0;
  if (LPN > 1) {
    if (MODE != 2) {
      double sums[4];
      node_sums_cm<D, LPN>(m, G, A, t, lane, sums);
#pragma unroll
      for (int i = 0; i < D; i++) f[i] = sums[i];
    }
    if (lane != 0 || !live) return;
  }
  for (int w = 0; LPN == 1 && MODE != 2 && w < G.w2t; w++) {
    uint32_t mm = G.occm[(size_t)w * G.max_act + t];
    while (mm) {
      const int q = w * 32 + __ffs(mm) - 1;
      mm &= mm - 1;
      const double* src = G.part + ((size_t)q * G.max_act + t) * D;
#pragma unroll
      for (int i = 0; i < D; i++) f[i] += src[i];
    }
  }
  if (MODE == 1 || (MODE == 3 && G.band[A])) {
#pragma unroll
    for (int i = 0; i < D; i++) G.F[(size_t)A * D + i] = f[i];
    return;
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
__global__ void __launch_bounds__(128, D == 2 ? 5 : 3) k_g2p(MeshDev m, PartDev P, GridDev G, StepParams sp, BlockCfg cfg, const AlmeDev al) {
  extern __shared__ __align__(16) unsigned char smem[];
  const LayoutC<D> L(cfg);
  int* s_rank = (int*)(smem + L.rank); double* s_X = (double*)(smem + L.X);
  double* s_U = (double*)(smem + L.U); double* s_A = (double*)(smem + L.A);
  double* s_tab = (double*)(smem + L.tab);
  if (threadIdx.x < 32) s_tab[threadIdx.x] = g_exp2tab[threadIdx.x];
  const int nocc = *G.n_occ, ngroups = (nocc + cfg.C - 1) / cfg.C;
  // persistent blocks: each walks over cell groups blockIdx.x, blockIdx.x + gridDim.x, ... (pipelined)
  NLPS_PIPE_BEGIN(false, 2, nullptr, s_U, s_A)
  const int SL = cfg.SL, np = P.ld;  // np: SoA stride
#ifndef NLPS_G2P_LP
#define NLPS_G2P_LP 1  // measured on the 64^3 cube: 1.11 ms with one lane per particle, 1.26 ms with two
#endif
  constexpr int LP = NLPS_G2P_LP > 1 ? LanesPerParticle<D, W>::value : 1;
  for (int jt = threadIdx.x; jt < (b.t1 - b.t0) * LP; jt += blockDim.x) {
    const int t = b.t0 + jt / LP, sub = jt % LP;
    const unsigned pm = (LP == 1) ? 0u : (((1u << LP) - 1u) << ((threadIdx.x & 31) & ~(LP - 1)));
    (void)pm;
    const int p = G.plist[t];
    const int ci = cell_of(s_cs, b.ncell, t);
    const double* Xc = s_X + (size_t)ci * SL * D;
    const double* Uc = s_U + (size_t)ci * SL * D;
    const double* Ac = s_A + (size_t)ci * SL * D;
    double xp[D], lam[D], a[D], du[D], pvl[D], pds[D];
#pragma unroll
    for (int i = 0; i < D; i++) {
      xp[i] = P.x[i * np + p]; lam[i] = P.lam[i * np + p]; a[i] = 0.0; du[i] = 0.0;
      pvl[i] = P.vel[i * np + p]; pds[i] = P.dis[i * np + p];  // consumed by the corrector after the loop
    }
    const double beta = P.beta[p];
    const MetricB MB = metric_load<D>(al, P.ld, p);  // aLME: the particle's metric instead of beta
    double Z = 0.0;
    uint32_t mk[W];
#pragma unroll
    for (int w = 0; w < W; w++) mk[w] = (LP == 1 || (w % LP) == sub) ? P.mask[(size_t)w * np + p] : 0u;
    for_slots<D, W>(mk, s_len[ci], [&](int k0, int k1, double w0, double w1) {
      double X0[D], X1[D], U0[D], U1[D], A0[D], A1[D], ll0 = 0.0, lx0 = 0.0, ll1 = 0.0, lx1 = 0.0;
      ldsvec<D>(Xc + k0 * D, X0);
      ldsvec<D>(Xc + k1 * D, X1);
      ldsvec<D>(Uc + k0 * D, U0);
      ldsvec<D>(Uc + k1 * D, U1);
      ldsvec<D>(Ac + k0 * D, A0);
      ldsvec<D>(Ac + k1 * D, A1);
      double l0[D], l1[D];
#pragma unroll
      for (int i = 0; i < D; i++) {
        l0[i] = xp[i] - X0[i];
        l1[i] = xp[i] - X1[i];
        ll0 += l0[i] * l0[i];
        ll1 += l1[i] * l1[i];
        lx0 += l0[i] * lam[i];
        lx1 += l1[i] * lam[i];
      }
      const double e0 = fexp(-metric_q<D>(MB, beta, ll0, l0) + lx0, s_tab) * w0;
      const double e1 = fexp(-metric_q<D>(MB, beta, ll1, l1) + lx1, s_tab) * w1;
      Z += e0;
#pragma unroll
      for (int i = 0; i < D; i++) {
        a[i] += e0 * A0[i];
        du[i] += e0 * U0[i];
      }
      Z += e1;
#pragma unroll
      for (int i = 0; i < D; i++) {
        a[i] += e1 * A1[i];
        du[i] += e1 * U1[i];
      }
    });
    if (LP > 1) {
      Z = pair_sum<LP>(Z, pm);
#pragma unroll
      for (int i = 0; i < D; i++) { a[i] = pair_sum<LP>(a[i], pm); du[i] = pair_sum<LP>(du[i], pm); }
      if (sub != 0) continue;
    }
    const double Zi = 1.0 / Z;
#pragma unroll
    for (int i = 0; i < D; i++) {
      const double ai = a[i] * Zi, di = du[i] * Zi;
      P.acc[i * np + p] = ai;
      P.ddis[i * np + p] = di;
      P.vel[i * np + p] = pvl[i] + 0.5 * sp.dt * ai;
      P.x[i * np + p] = xp[i] + di;
      P.dis[i * np + p] = pds[i] + di;
    }
  }
  __syncthreads();
  NLPS_PIPE_END()  // cell groups
}


// Reference roll semantics for history variables of particles whose law never writes them
// (Neo-Hookean: b_e, EPS, Kappa): U-Verlet.c:1043-1056 COPIES n1 -> n every step, so after the
// first step both copies hold the initial n1 value.  With the pointer-swap roll that is reproduced by
// copying n1 -> n once, before the first swap.
template <int D>
__global__ void k_sync_inert(PartDev P, const __grid_constant__ MatTable mt) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.np) return;
  const int mtype = mt.m[P.matidx[p]].type;
  if (mtype == NLPS_MAT_VON_MISES) P.kap_n[p] = P.kap_n1[p];  // Von-Mises updates b_e and EPS but never Kappa
  if (mat_has_history(mtype)) return;
  constexpr int TB = (D == 2) ? 5 : 9;
#pragma unroll
  for (int i = 0; i < TB; i++) P.be_n[(size_t)i * P.ld + p] = P.be_n1[(size_t)i * P.ld + p];
  P.eps_n[p] = P.eps_n1[p];
  P.kap_n[p] = P.kap_n1[p];
}

// ---------------------------------------------------------------------------
// Slab halo (multi-GPU): every slab keeps the nodes within a band of its cuts; the sums of
// the shared nodes are exchanged with the neighbour slab and added.  which: 0 = cell occupancy (-> remote
// occupancy flags, feeds ActiveNode), 1 = lumped mass + momentum sums, 2 = force sums.
template <int D>
__global__ void k_halo_pack(GridDev G, const int* ids, int n, int which, double* buf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int A = ids[i];
  if (which == 0) {
    buf[i] = G.cnt[A] > 0 ? 1.0 : 0.0;
  } else if (which == 1) {
    const bool act = G.active[A];
    buf[(size_t)i * (1 + D)] = act ? G.M[A] : 0.0;
#pragma unroll
    for (int k = 0; k < D; k++) buf[(size_t)i * (1 + D) + 1 + k] = act ? G.MOM[(size_t)A * D + k] : 0.0;
  } else {
    const bool act = G.active[A];
#pragma unroll
    for (int k = 0; k < D; k++) buf[(size_t)i * D + k] = act ? G.F[(size_t)A * D + k] : 0.0;
  }
}
template <int D>
__global__ void k_halo_add(GridDev G, const int* ids, int n, int which, const double* buf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int A = ids[i];
  if (which == 0) {
    G.rocc[A] = (unsigned char)(buf[i] != 0.0);  // the bands of the two cuts are disjoint: plain overwrite
    if (buf[i] != 0.0) G.occ_blk[A >> 8] = 1;
  } else if (which == 1) {
    if (!G.active[A]) return;
    G.M[A] += buf[(size_t)i * (1 + D)];
#pragma unroll
    for (int k = 0; k < D; k++) G.MOM[(size_t)A * D + k] += buf[(size_t)i * (1 + D) + 1 + k];
  } else {
    if (!G.active[A]) return;
#pragma unroll
    for (int k = 0; k < D; k++) G.F[(size_t)A * D + k] += buf[(size_t)i * D + k];
  }
}

// Peer-memory halo exchange (one process per GPU, neighbours' buffers mapped with CUDA IPC): the push kernel packs
// the band values of BOTH sides and stores them straight into the neighbours' receive buffers over NVLink, the last
// block to finish publishes the arrival counters; the pull kernel waits for the neighbours' counters and adds.
// A push never waits, so the pair cannot deadlock; the wait is bounded (30 s, NLPS_HALO_TIMEOUT_S) and latches
// NLPS_ERR_HALO_TIMEOUT.  The receive buffers are doubled by the parity of the exchange number.
struct HaloP2P {
  const int* ids[2];
  int n[2];
  double* peer_rbuf[2];               // where my values go (offset for this exchange kind already applied)
  unsigned long long* peer_flag[2];   // the neighbour's counter for this kind
  const double* my_rbuf[2];
  const unsigned long long* my_flag[2];
};
template <int D>
__device__ __forceinline__ void halo_values(const GridDev& G, int A, int which, double* v) {
  if (which == 0) {
    v[0] = G.cnt[A] > 0 ? 1.0 : 0.0;
  } else if (which == 1) {
    const bool act = G.active[A];
    v[0] = act ? G.M[A] : 0.0;
#pragma unroll
    for (int k = 0; k < D; k++) v[1 + k] = act ? G.MOM[(size_t)A * D + k] : 0.0;
  } else {
    const bool act = G.active[A];
#pragma unroll
    for (int k = 0; k < D; k++) v[k] = act ? G.F[(size_t)A * D + k] : 0.0;
  }
}
template <int D>
__global__ void __launch_bounds__(256) k_halo_push(GridDev G, HaloP2P h, int which, unsigned long long seq, unsigned int* done) {
  const int per = (which == 0) ? 1 : ((which == 1) ? 1 + D : D);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int s_ = (i < h.n[0]) ? 0 : 1, j = (s_ == 0) ? i : i - h.n[0];
  if (j < h.n[s_]) {
    double v[1 + D];
    halo_values<D>(G, h.ids[s_][j], which, v);
    for (int k = 0; k < per; k++) h.peer_rbuf[s_][(size_t)j * per + k] = v[k];
  }
  __threadfence_system();  // my stores are visible to the peer before the counter can be
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(done, 1u);
    if (ticket == gridDim.x - 1) {
      *done = 0u;
      __threadfence_system();
      for (int t = 0; t < 2; t++)
        if (h.n[t] > 0) *((volatile unsigned long long*)h.peer_flag[t]) = seq;
    }
  }
}
template <int D>
__global__ void __launch_bounds__(256) k_halo_pull(GridDev G, HaloP2P h, int which, unsigned long long seq, int* err,
                                                   long long timeout_cycles) {
  const int per = (which == 0) ? 1 : ((which == 1) ? 1 + D : D);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int s_ = (i < h.n[0]) ? 0 : 1, j = (s_ == 0) ? i : i - h.n[0];
  if (j >= h.n[s_]) return;
  const volatile unsigned long long* f = (const volatile unsigned long long*)h.my_flag[s_];
  const long long t0 = clock64();
  while (*f < seq) {
    if (clock64() - t0 > timeout_cycles) { latch_error(err, NLPS_ERR_HALO_TIMEOUT, -1); return; }
    __nanosleep(64);
  }
  __threadfence_system();
  const double* buf = h.my_rbuf[s_] + (size_t)j * per;
  const int A = h.ids[s_][j];
  if (which == 0) {
    const bool on = ((const volatile double*)buf)[0] != 0.0;
    G.rocc[A] = (unsigned char)on;
    if (on) G.occ_blk[A >> 8] = 1;
  } else if (which == 1) {
    if (!G.active[A]) return;
    G.M[A] += ((const volatile double*)buf)[0];
#pragma unroll
    for (int k = 0; k < D; k++) G.MOM[(size_t)A * D + k] += ((const volatile double*)buf)[1 + k];
  } else {
    if (!G.active[A]) return;
#pragma unroll
    for (int k = 0; k < D; k++) G.F[(size_t)A * D + k] += ((const volatile double*)buf)[k];
  }
}

// ---------------------------------------------------------------------------
// Migration between slabs: particles whose closest node crossed a cut move to the neighbour.
template <int D>
__global__ void __launch_bounds__(256) k_mig_mark(MeshDev m, PartDev P, SlabDev sl, int* dest, int* cnt) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  int d = -1;
  if (p < P.np) {
    const double c = m.X[(size_t)P.I0[p] * NS<D>::X + sl.axis];
    d = (c < sl.own_lo) ? 1 : ((c >= sl.own_hi) ? 2 : 0);
    dest[p] = d;
  }
  const int n0 = __syncthreads_count(d == 0), n1 = __syncthreads_count(d == 1), n2 = __syncthreads_count(d == 2);
  if (threadIdx.x == 0) {
    if (n0) atomicAdd(&cnt[0], n0);
    if (n1) atomicAdd(&cnt[1], n1);
    if (n2) atomicAdd(&cnt[2], n2);
  }
}
// permutation: stayers first, then the particles bound for the lower slab, then for the upper slab
__global__ void __launch_bounds__(256) k_mig_perm(int np, const int* dest, int* cnt, int* perm) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= np) return;
  const int d = dest[p];
  const int base = (d == 0) ? 0 : ((d == 1) ? cnt[0] : cnt[0] + cnt[1]);
  perm[base + atomicAdd(&cnt[3 + d], 1)] = p;
}
struct MigCol { void* base; int is_int; int pad; unsigned long long off; };  // off: byte offset of the column per buffered row
// rows [row0, row0 + nrows) of every column <-> buffer (column c at buf + off_c * nrows)
// (is_int == 2: a node id -- travels as a GLOBAL id: + node_offset on the way out, - node_offset on the way in)
__global__ void __launch_bounds__(256) k_mig_copy(const MigCol* tab, int ncols, int row0, int nrows, unsigned char* buf, int unpack,
                                                  int node_offset) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)ncols * nrows) return;
  const int c = (int)(i / nrows), r = (int)(i % nrows);
  const MigCol col = tab[c];
  if (col.is_int) {
    int* a = (int*)col.base + row0 + r;
    int* b = (int*)(buf + col.off * nrows) + r;
    const int shift = (col.is_int == 2) ? node_offset : 0;
    if (unpack) *a = *b - shift; else *b = *a + shift;
  } else {
    double* a = (double*)col.base + row0 + r;
    double* b = (double*)(buf + col.off * nrows) + r;
    if (unpack) *a = *b; else *b = *a;
  }
}
__global__ void __launch_bounds__(256) k_mig_inv(PartDev P, int row0, int nrows, int value_is_slot) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  P.inv[P.orig[row0 + r]] = value_is_slot ? row0 + r : -1;
}
__global__ void k_fill_i(int* a, size_t n, int v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}
__global__ void k_set_inv(const int* orig, int* inv, int n) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) inv[orig[p]] = p;
}

// AoS (host layout, rows x cols) <-> SoA (cols x ld)
// (row rowmap[p] of the AoS buffer <-> physical slot p; rowmap == nullptr: row p)
__global__ void k_aos_to_soa(const double* aos, double* soa, const int* rowmap, int n, int ld, int cols, int aos_stride, int col0) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * cols) return;
  int c = (int)(i / n), p = (int)(i % n);
  const int row = rowmap ? rowmap[p] : p;
  soa[(size_t)c * ld + p] = aos[(size_t)row * aos_stride + col0 + c];
}
__global__ void k_soa_to_aos(const double* soa, double* aos, const int* rowmap, int n, int ld, int cols, int aos_stride, int col0) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * cols) return;
  int p = (int)(i / cols), c = (int)(i % cols);
  const int row = rowmap ? rowmap[p] : p;
  aos[(size_t)row * aos_stride + col0 + c] = soa[(size_t)c * ld + p];
}
__global__ void k_unpermute_int(const int* src, int* dst, const int* rowmap, int n) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) dst[rowmap ? rowmap[p] : p] = src[p];
}
__global__ void k_permute_int(const int* src, int* dst, const int* rowmap, int n) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) dst[p] = src[rowmap ? rowmap[p] : p];
}
__global__ void k_iota(int* a, int* b, int n) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) { a[p] = p; b[p] = p; }
}
__global__ void k_fill_d(double* a, size_t n, double v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}
// expand bitmask lists to the reference's ListNodes order (reverse of acceptance order)
__global__ void k_expand_lists(MeshDev m, PartDev P, int W, int cap, int* lists, int compact) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.np) return;
  int base = m.r2p[P.I0[p]], o = 0;
  const size_t row = compact ? p : P.orig[p];
  for (int w = W - 1; w >= 0; w--) {
    uint32_t mm = P.mask[(size_t)w * P.ld + p];
    while (mm) {
      int b = 31 - __clz(mm);
      mm &= ~(1u << b);
      if (o < cap) lists[row * cap + o] = m.r2i[base + w * 32 + b];
      o++;
    }
  }
  for (; o < cap; o++) lists[row * cap + o] = -1;
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
__global__ void k_stress_points(int n, const __grid_constant__ MatTable mt, ReturnMapParams rp, const double* DF, const double* F1, const double* J1,
                                const double* be_n, const double* eps_n, const double* kap_n, double* stress,
                                double* be_n1, double* eps_n1, double* kap_n1, double* W, double* cep, int* status) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  constexpr int T = (D == 2) ? 5 : 9;
  const MatParams& m = mt.m[0];
  double df[D * D], f1[D * D], be[T], tau[T], c[D * D], Wp = 0.0;
#pragma unroll
  for (int i = 0; i < D * D; i++) { df[i] = DF[(size_t)p * T + i]; f1[i] = F1[(size_t)p * T + i]; c[i] = 0.0; }
#pragma unroll
  for (int i = 0; i < T; i++) be[i] = be_n[(size_t)p * T + i];
  double eps = eps_n[p], kap = kap_n[p];
  int st = 0;
  if (m.type == NLPS_MAT_NEO_HOOKEAN_WRIGGERS) stress_neo_hookean<D>(m, f1, J1[p], tau, Wp);
  else {
    double back[3] = {0.0, 0.0, 0.0};  // material points carry no back stress: Von-Mises starts from zero
    st = stress_with_history<D>(m.type, m, rp, df, f1, be, eps, kap, back, tau, Wp, c);
  }
  status[p] = st;
#pragma unroll
  for (int i = 0; i < T; i++) { stress[(size_t)p * T + i] = tau[i]; be_n1[(size_t)p * T + i] = be[i]; }
  eps_n1[p] = eps; kap_n1[p] = kap; W[p] = Wp;
#pragma unroll
  for (int i = 0; i < D * D; i++) cep[(size_t)p * D * D + i] = c[i];
}

// ---------------------------------------------------------------------------
// Host-side engine
struct ImplicitCtx;
static void implicit_free(nlps_engine* e);
struct nlps_engine {
  int D = 2, T = 5, TB = 5, W = 1, np = 0, nn = 0, device = 0, cap = 0;
  nlps_solver solver{};
  cudaStream_t stream = nullptr;
  MeshDev mesh{};
  PartDev P{};
  AlmeDev alme{};  // aLME shape functions (2D): metric and cut-off ellipsoid per particle, else nullptr
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
  MatTable mat{};        // this engine's material table (kernel parameter of the stress kernels)
  // kernel generation of the three hot stages: 2 = warp-per-cell kernels (nlps_cellwarp.cu), 1 = block-per-cell-group
  // kernels (below).  NLPS_KERNELS overrides; NLPS_SPLIT_NH=1 runs Neo-Hookean clouds through gather / stress / force too
  int kver = 2, split_nh = 0;
  CwCfg cw{};
  CwState cws{};
  int inert_synced = 0;
  BlockCfg cfg{};          // cells per block / 2-ring row length / particle chunk of the cell-block kernels
  size_t smemA = 0, smemB = 0, smemC = 0;
  int cache_pa = 1;        // keep the shape-function weights of the particle phase in shared memory (2D)
  int smem_set[K_COUNT] = {0};
  int reorder_every = 25;  // physical cell-sort cadence (steps); the results do not depend on it
  int steps_since_sort = 1 << 30;
  long long n_reorders = 0;
  int max_smem_optin = 0;
  int sm_count = 0, grid_override = 0;
  int grid_k[K_COUNT] = {0};  // blocks of the persistent cell-group kernels (SMs x resident blocks)
  std::vector<void*> allocs;
  // staging for AoS <-> SoA
  double* stage = nullptr;
  size_t stage_doubles = 0;
  // asynchronous results download (scheme call): fields are transposed into `snap` on the compute stream and copied to
  // the caller's buffers on `dl_stream` while the next steps run
  double* snap = nullptr;
  size_t snap_doubles = 0, snap_off = 0;
  int dl_async = 0, dl_pending = 0;
  cudaStream_t dl_stream = nullptr;
  cudaEvent_t dl_ready = nullptr, dl_done = nullptr;
  struct DlCopy { void* h; const void* d; size_t bytes; };
  std::vector<DlCopy> dl_copies;
  // profiling
  int profile = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double k_ms[K_COUNT] = {0};
  int k_n[K_COUNT] = {0};
  long long launches = 0;
  int last_code = 0, last_particle = -1;
  int host_fail = 0;  // a host-side failure (migration, halo transport): sticky, reported by every later poll
  // The n+1 -> n roll is a pointer swap: after a completed step the device *_n1 arrays hold the values of step n - 1,
  // while the reference COPIES n1 -> n (U-Verlet.c:1040-1081) and leaves both equal.  Downloads taken in that state
  // write the *_n arrays into both host fields; inside a step (after the kinematics stage) n1 is what it says.
  int n1_stale = 0;
  // ---- spatial slab (multi-GPU); slab_on == 0: the engine owns every particle
  int slab_on = 0, rank = 0, world = 1, axis = 0, band_cells = 6, migrate_every = 10, n_global = 0;
  int steps_since_migration = 0;
  long long n_migrated_in = 0;
  double cut_lo = -1e300, cut_hi = 1e300;  // ownership interval of closest-node coordinates [cut_lo, cut_hi)
  nlps_comm* comm = nullptr;
  struct HaloSide {
    int peer = -1, n = 0;
    int* ids = nullptr;            // device, ascending node ids, identical on both sides of the cut
    double *sbuf = nullptr, *rbuf = nullptr;  // n x (1 + D)
    // peer-memory path (NVLink stores into the neighbour's buffers, no NCCL launch): 3 receive buffers (one per
    // exchange kind) + 3 arrival counters in MY memory, and the neighbour's, mapped through CUDA IPC
    double* p2p_rbuf = nullptr;               // mine: 3 x n x (1 + D), cudaMalloc (IPC needs it)
    unsigned long long* p2p_flag = nullptr;   // mine: [4]
    double* peer_rbuf = nullptr;              // the neighbour's p2p_rbuf (of its side facing me)
    unsigned long long* peer_flag = nullptr;
  } side[2];                        // 0: lower neighbour, 1: upper neighbour
  int p2p_on = 0;
  unsigned long long halo_seq[3] = {0, 0, 0};
  unsigned int* p2p_done = nullptr;  // device counter of the push kernel's last-block election
  int* mig_dest = nullptr;          // per particle: 0 stay, 1 to the lower slab, 2 to the upper slab
  int* mig_cnt = nullptr;           // device [8]: counts stay/low/up, cursors, received low/up
  int* h_mig = nullptr;             // pinned [8]
  unsigned char *mig_sbuf[2] = {nullptr, nullptr}, *mig_rbuf[2] = {nullptr, nullptr};
  int mig_cap = 0;                  // rows per migration buffer
  MigCol* mig_tab = nullptr;
  double solver_dx = 0.0;           // Mesh.DeltaX
  int node_offset = 0;              // global node id = local id + node_offset (sub-mesh slabs)
  int implicit_on = 0;              // inside an implicit (Newmark-beta) step: kinematics leave rho alone
  struct ImplicitCtx* imp = nullptr;
  std::vector<int> h_ids;           // host copy of P.orig (slab I/O)
  std::vector<double> h_rows;       // host staging of compact rows (slab I/O)
  void* p2p_ce[2] = {nullptr, nullptr};  // P2PCacheEntry of each side while this engine uses it
  int* d_rows = nullptr;            // create: state-row index of every held particle (device-side row gather)
  int up_rows = 0;                  // create: rows of the caller's state buffers
};

// ---------------------------------------------------------------------------
// NCCL is bound with dlopen at the first use, not at link time: a process that also hosts PyTorch must end
// up with ONE libnccl.so.2 (the dynamic linker deduplicates by soname), and it has to be the newer one
// PyTorch ships; linking would pin the system copy as soon as this library is loaded.
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi* nccl_api() {
  static NcclApi api;
  static int tried = 0;
  if (!tried) {
    tried = 1;
    const char* names[] = {getenv("NLPS_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm) continue;
      api.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
    }
    if (api.h) {
#define SYM_(f) *(void**)(&api.f) = dlsym(api.h, "nccl" #f)
      SYM_(GetUniqueId); SYM_(CommInitRank); SYM_(CommDestroy); SYM_(GroupStart); SYM_(GroupEnd); SYM_(Send); SYM_(Recv);
      SYM_(AllReduce); SYM_(GetErrorString);
#undef SYM_
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.GroupStart || !api.GroupEnd || !api.Send ||
          !api.Recv || !api.AllReduce || !api.GetErrorString)
        api.h = nullptr;
    }
    if (!api.h) fprintf(stderr, "nlps_b200: libnccl.so.2 not found (set NLPS_NCCL_LIB)\n");
  }
  return api.h ? &api : nullptr;
}

// Transports of the slab exchanges
struct nlps_comm {
  int rank = 0, world = 1, is_nccl = 0;
  ncclComm_t nccl = nullptr;
  nlps_exchange_fn fn = nullptr;
  void* user = nullptr;
};
static int comm_exchange(nlps_comm* c, int n, const nlps_msg* msgs, cudaStream_t stream) {
  if (!c) return n ? 1 : 0;
  if (!c->is_nccl) return c->fn(c->user, n, msgs, (void*)stream);  // also with n == 0: collective transports count calls
  if (n == 0) return 0;
  NcclApi* N = nccl_api();
  ncclResult_t r = N->GroupStart();
  for (int i = 0; i < n && r == ncclSuccess; i++) {
    if (msgs[i].send_bytes) r = N->Send(msgs[i].send, msgs[i].send_bytes, ncclChar, msgs[i].peer, c->nccl, stream);
    if (r == ncclSuccess && msgs[i].recv_bytes) r = N->Recv(msgs[i].recv, msgs[i].recv_bytes, ncclChar, msgs[i].peer, c->nccl, stream);
  }
  ncclResult_t r2 = N->GroupEnd();
  if (r != ncclSuccess || r2 != ncclSuccess) {
    fprintf(stderr, "nlps_b200: NCCL exchange failed: %s\n", N->GetErrorString(r != ncclSuccess ? r : r2));
    return 1;
  }
  return 0;
}

// Collective "did everybody succeed": minimum of `ok` (0 / 1) over all slabs.  NCCL: one ncclAllReduce on the engine's
// stream.  Custom transports only know neighbour exchanges: the minimum travels along the chain of slabs in world - 1
// rounds.  d_buf: device int[4] (value | to neighbours | from lower | from upper).  Every slab must call it.
static int comm_all_ok(nlps_comm* c, int rank, int world, int ok, int* d_buf, cudaStream_t stream, int* verdict) {
  *verdict = ok;
  if (!c || world <= 1) return 0;
  int v = ok ? 1 : 0;
  if (c->is_nccl) {
    NcclApi* N = nccl_api();
    if (cudaMemcpyAsync(d_buf, &v, sizeof(int), cudaMemcpyHostToDevice, stream) != cudaSuccess) return 1;
    if (N->AllReduce(d_buf, d_buf, 1, ncclInt, ncclMin, c->nccl, stream) != ncclSuccess) return 1;
    if (cudaMemcpyAsync(&v, d_buf, sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess)
      return 1;
    *verdict = v;
    return 0;
  }
  for (int round = 0; round < world - 1; round++) {
    int got[2] = {1, 1};
    if (cudaMemcpyAsync(d_buf, &v, sizeof(int), cudaMemcpyHostToDevice, stream) != cudaSuccess) return 1;
    nlps_msg msgs[2];
    int nm = 0;
    if (rank > 0) msgs[nm++] = nlps_msg{rank - 1, d_buf, sizeof(int), d_buf + 2, sizeof(int)};
    if (rank < world - 1) msgs[nm++] = nlps_msg{rank + 1, d_buf, sizeof(int), d_buf + 3, sizeof(int)};
    if (comm_exchange(c, nm, msgs, stream)) return 1;
    if (cudaMemcpyAsync(got, d_buf + 2, sizeof(got), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess)
      return 1;
    if (rank > 0) v = std::min(v, got[0]);
    if (rank < world - 1) v = std::min(v, got[1]);
  }
  *verdict = v;
  return 0;
}


// Element-wise sum of a short device vector over all slabs, result identical (bit for bit) on every slab.  NCCL: one
// ncclAllReduce on the engine's stream, no host round trip.  Custom transports only know neighbour exchanges: every
// slab's vector travels along the chain (world - 1 rounds, both directions) into a table tab[(world + 1) x n] (last row =
// scratch for the messages that have nothing to carry) and the rows are added in rank order.
__global__ void k_sum_rows(const double* tab, int rows, int n, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s_ = 0.0;
  for (int r = 0; r < rows; r++) s_ += tab[(size_t)r * n + i];
  out[i] = s_;
}
static int comm_allreduce_sum(nlps_comm* c, int rank, int world, double* d_vec, int n, double* tab, cudaStream_t stream) {
  if (!c || world <= 1) return 0;
  if (c->is_nccl) return nccl_api()->AllReduce(d_vec, d_vec, (size_t)n, ncclDouble, ncclSum, c->nccl, stream) == ncclSuccess ? 0 : 1;
  auto row = [&](int r) { return tab + (size_t)((r >= 0 && r < world) ? r : world) * n; };
  if (cudaMemcpyAsync(row(rank), d_vec, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream) != cudaSuccess) return 1;
  const unsigned long long bytes = sizeof(double) * (unsigned long long)n;
  for (int r = 1; r < world; r++) {
    nlps_msg msgs[2];
    int nm = 0;
    // to the upper neighbour goes the row that came from below in the previous round (my own in the first), and back
    if (rank > 0) msgs[nm++] = nlps_msg{rank - 1, row(rank + r - 1), bytes, row(rank - r), bytes};
    if (rank < world - 1) msgs[nm++] = nlps_msg{rank + 1, row(rank - r + 1), bytes, row(rank + r), bytes};
    if (comm_exchange(c, nm, msgs, stream)) return 1;
  }
  k_sum_rows<<<(n + 255) / 256, 256, 0, stream>>>(tab, world, n, d_vec);
  return 0;
}

// Device memory comes from the device's stream-ordered pool with an unlimited release threshold: an engine
// destroyed and re-created in the same process (a second scheme call, a parameter sweep) gets its blocks back
// from the pool instead of paying cudaMalloc / cudaFree again (0.8 s + 0.8 s for the 10^6-particle deck).
// nlps_b200_trim() hands the cached memory back; NLPS_POOL=0 falls back to cudaMalloc / cudaFree.
static bool use_pool() {
  static int v = -1;
  if (v < 0) {
    const char* s_ = getenv("NLPS_POOL");
    v = (s_ && atoi(s_) == 0) ? 0 : 1;
  }
  return v == 1;
}
static void pool_setup(int device) {
  if (!use_pool()) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
}
static cudaError_t pool_malloc(void** q, size_t bytes, cudaStream_t stream) {
  return use_pool() ? cudaMallocAsync(q, bytes, stream) : cudaMalloc(q, bytes);
}
static void pool_free(void* q, cudaStream_t stream) {
  if (use_pool()) cudaFreeAsync(q, stream); else cudaFree(q);
}

template <typename Tp>
static int dev_alloc(nlps_engine* e, Tp** p, size_t n) {
  void* q = nullptr;
  cudaError_t st = pool_malloc(&q, std::max<size_t>(n, 1) * sizeof(Tp), e->stream);
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

// transposed adjacency (who lists me), rows in ascending source order; qpos[q] = position of entry q inside the
// transposed row of idx[q].  Threads own disjoint ranges of DESTINATION nodes and each scans the whole
// adjacency: no atomics, the same order as the serial algorithm whatever the thread count.
static void transpose_csr(int nn, const int* ptr, const int* idx, std::vector<int>& tp, std::vector<int>& ti,
                          std::vector<unsigned char>* qpos = nullptr) {
  tp.assign(nn + 1, 0);
  const long long nnz = ptr[nn];
  if (qpos) qpos->resize(nnz);
#pragma omp parallel
  {
    int nt = 1, me = 0;
#ifdef _OPENMP
    nt = omp_get_num_threads();
    me = omp_get_thread_num();
#endif
    const int a0 = (int)((long long)nn * me / nt), a1 = (int)((long long)nn * (me + 1) / nt);
    for (long long q = 0; q < nnz; q++) {
      const int A = idx[q];
      if (A >= a0 && A < a1) tp[A + 1]++;
    }
#pragma omp barrier
#pragma omp single
    {
      for (int i = 0; i < nn; i++) tp[i + 1] += tp[i];
      ti.resize(tp[nn]);
    }
    std::vector<int> fill(tp.begin() + a0, tp.begin() + a1);
    for (int i = 0; i < nn; i++)
      for (int q = ptr[i]; q < ptr[i + 1]; q++) {
        const int A = idx[q];
        if (A < a0 || A >= a1) continue;
        int& f = fill[A - a0];
        if (qpos) (*qpos)[q] = (unsigned char)(f - tp[A]);
        ti[f++] = i;
      }
  }
}

// ---- the same transposition on the device (create-time; 0.3 s of host time for the 10^6-particle deck otherwise):
// histogram, exclusive scan, unordered fill, per-row sort (rows hold at most 255 sources) -> ascending source order,
// i.e. exactly what transpose_csr() produces
__global__ void __launch_bounds__(256) k_tr_count(const int* idx, long long nnz, int* cnt) {
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nnz) atomicAdd(&cnt[idx[q]], 1);
}
__global__ void __launch_bounds__(256) k_tr_pack(const int* cnt, ulonglong2* packed, int nn) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= nn) packed[i] = make_ulonglong2(i < nn ? (unsigned long long)cnt[i] : 0ull, 0ull);
}
__global__ void __launch_bounds__(256) k_tr_fill(const int* ptr, const int* idx, const int* tp, int* cursor, int* ti, int nn) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  for (int q = ptr[i]; q < ptr[i + 1]; q++) {
    const int A = idx[q];
    ti[tp[A] + atomicAdd(&cursor[A], 1)] = i;
  }
}
__global__ void __launch_bounds__(256) k_tr_sort(const int* tp, int* ti, int nn) {
  int A = blockIdx.x * blockDim.x + threadIdx.x;
  if (A >= nn) return;
  int* a = ti + tp[A];
  const int n = tp[A + 1] - tp[A];
  for (int i = 1; i < n; i++) {
    const int v = a[i];
    int j = i - 1;
    while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; j--; }
    a[j + 1] = v;
  }
}
__global__ void __launch_bounds__(256) k_tr_qpos(const int* ptr, const int* idx, const int* tp, const int* ti, unsigned char* qpos, int nn) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  for (int q = ptr[i]; q < ptr[i + 1]; q++) {
    const int A = idx[q];
    int lo = tp[A], hi = tp[A + 1];
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (ti[mid] < i) lo = mid + 1; else hi = mid;
    }
    qpos[q] = (unsigned char)(lo - tp[A]);
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

// persistent cell-group kernels: dynamic shared memory opt-in and grid = SMs x resident blocks, both
// resolved at the first launch of the instantiation this engine uses
#define LAUNCH_SMEM(e, id, kernel, max_blocks, block, smem_bytes, ...)                           \
  do {                                                                                          \
    if (!(e)->smem_set[id]) {                                                                   \
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (e)->max_smem_optin); \
      int _nb = 1;                                                                              \
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&_nb, kernel, (block), (smem_bytes));       \
      (e)->grid_k[id] = (e)->sm_count * std::max(1, _nb);                                       \
      if ((e)->grid_override > 0) (e)->grid_k[id] = (e)->grid_override;                         \
      (e)->smem_set[id] = 1;                                                                    \
    }                                                                                           \
    if ((e)->profile) cudaEventRecord((e)->ev0, (e)->stream);                                   \
    kernel<<<std::min((max_blocks), (e)->grid_k[id]), (block), (smem_bytes), (e)->stream>>>(__VA_ARGS__); \
    (e)->launches++;                                                                            \
    if ((e)->profile) {                                                                         \
      cudaEventRecord((e)->ev1, (e)->stream);                                                   \
      cudaEventSynchronize((e)->ev1);                                                           \
      float _ms = 0;                                                                            \
      cudaEventElapsedTime(&_ms, (e)->ev0, (e)->ev1);                                           \
      (e)->k_ms[id] += _ms;                                                                     \
      (e)->k_n[id]++;                                                                           \
    }                                                                                           \
  } while (0)

static inline int nblk(size_t n, int b) { return (int)((n + b - 1) / b); }

// device transposition of a CSR adjacency already uploaded (d_ptr, d_idx); returns the longest transposed row
static int device_transpose(nlps_engine* e, int nn, const int* d_ptr, const int* d_idx, long long nnz, int** d_tp, int** d_ti,
                            unsigned char** d_qpos, int* max_row) {
  int *cnt = nullptr, *cursor = nullptr, *dummy_a = nullptr, *dummy_b = nullptr, *tops = nullptr;
  ulonglong2 *packed = nullptr, *blk = nullptr;
  const int items = nn + 1, nb = nblk(items, SCAN_ITEMS);
  if (dev_alloc(e, d_tp, (size_t)nn + 1) || dev_alloc(e, d_ti, (size_t)std::max<long long>(nnz, 1))) return 1;
  if (d_qpos && dev_alloc(e, d_qpos, (size_t)std::max<long long>(nnz, 1))) return 1;
  void* tmp[7] = {nullptr};
  auto talloc = [&](void** q, size_t bytes) {
    if (pool_malloc(q, std::max<size_t>(bytes, 16), e->stream) != cudaSuccess) return 1;
    cudaMemsetAsync(*q, 0, std::max<size_t>(bytes, 16), e->stream);
    return 0;
  };
  if (talloc(&tmp[0], sizeof(int) * nn) || talloc(&tmp[1], sizeof(int) * nn) || talloc(&tmp[2], sizeof(int) * items) ||
      talloc(&tmp[3], sizeof(int) * items) || talloc(&tmp[4], sizeof(int) * 4) || talloc(&tmp[5], sizeof(ulonglong2) * items) ||
      talloc(&tmp[6], sizeof(ulonglong2) * (nb + 1)))
    return 1;
  cnt = (int*)tmp[0]; cursor = (int*)tmp[1]; dummy_a = (int*)tmp[2]; dummy_b = (int*)tmp[3]; tops = (int*)tmp[4];
  packed = (ulonglong2*)tmp[5]; blk = (ulonglong2*)tmp[6];
  if (nnz) k_tr_count<<<nblk((size_t)nnz, 256), 256, 0, e->stream>>>(d_idx, nnz, cnt);
  k_tr_pack<<<nblk(items, 256), 256, 0, e->stream>>>(cnt, packed, nn);
  k_scan_reduce<<<nb, 256, 0, e->stream>>>(packed, blk, items);
  k_scan_tops<<<1, 1024, 0, e->stream>>>(blk, nb, tops, tops + 1, tops + 2);
  k_scan_apply<<<nb, 256, 0, e->stream>>>(packed, blk, *d_tp, dummy_a, dummy_b, items);
  k_tr_fill<<<nblk(nn, 256), 256, 0, e->stream>>>(d_ptr, d_idx, *d_tp, cursor, *d_ti, nn);
  k_tr_sort<<<nblk(nn, 256), 256, 0, e->stream>>>(*d_tp, *d_ti, nn);
  if (d_qpos) k_tr_qpos<<<nblk(nn, 256), 256, 0, e->stream>>>(d_ptr, d_idx, *d_tp, *d_ti, *d_qpos, nn);
  std::vector<int> h(nn);
  CUDA_OK(cudaMemcpyAsync(h.data(), cnt, sizeof(int) * nn, cudaMemcpyDeviceToHost, e->stream));
  CUDA_OK(cudaStreamSynchronize(e->stream));
  int m = 0;
  for (int i = 0; i < nn; i++) m = std::max(m, h[i]);
  *max_row = m;
  for (void* q : tmp) pool_free(q, e->stream);
  return 0;
}

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
  sp.proj = nullptr;
  sp.reuse_lists = 0;
  sp.implicit = 0;
  return sp;
}

// host copy of the global ids of the particles in slot order (slab I/O)
static int refresh_ids(nlps_engine* e) {
  e->h_ids.resize(e->np);
  if (e->np) CUDA_OK(cudaMemcpyAsync(e->h_ids.data(), e->P.orig, sizeof(int) * e->np, cudaMemcpyDeviceToHost, e->stream));
  CUDA_OK(cudaStreamSynchronize(e->stream));
  return 0;
}
// AoS host -> SoA device for one field (cols columns starting at col0 of an aos_stride-wide row).
// rows: how the host buffer is indexed -- 0: by the caller's particle id (whole buffer copied, rows picked on the
// device), 1: by global id, gathered on the host (slab engines hold a subset), 2: compact, row = slot.
static int put_field(nlps_engine* e, const double* h, double* d, int cols, int aos_stride, int col0, int rows = 0) {
  if (!h || !d || e->np == 0) return 0;
  size_t n = (size_t)e->np * aos_stride;
  const int* rowmap = e->P.orig;
  if (rows == 3) {  // the caller's whole buffer goes up as it is (pinned buffers: full link speed), rows picked on the device
    n = (size_t)e->up_rows * aos_stride;
    rowmap = e->d_rows;
  } else
  if (rows == 1) {
    e->h_rows.resize(n);
    for (int p = 0; p < e->np; p++) memcpy(&e->h_rows[(size_t)p * aos_stride], h + (size_t)e->h_ids[p] * aos_stride, sizeof(double) * aos_stride);
    h = e->h_rows.data();
    rowmap = nullptr;
  } else if (rows == 2) {
    rowmap = nullptr;
  }
  CUDA_OK(cudaMemcpyAsync(e->stage, h, n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  k_aos_to_soa<<<nblk((size_t)e->np * cols, 256), 256, 0, e->stream>>>(e->stage, d, rowmap, e->np, e->P.ld, cols, aos_stride, col0);
  CUDA_OK(cudaStreamSynchronize(e->stream));  // host buffer may be pageable; stage is reused (a pinned bounce buffer measured slower)
  return 0;
}
static int get_field(nlps_engine* e, double* h, const double* d, int cols, int aos_stride, int col0,
                     const double* d_extra = nullptr, int rows = 0) {
  if (!h || !d || e->np == 0) return 0;
  const size_t n = (size_t)e->np * aos_stride;
  const int* rowmap = rows == 0 ? e->P.orig : nullptr;
  double* stage = e->stage;
  if (e->dl_async && rows != 1) {  // own region of the snapshot, copied later on dl_stream (download_flush)
    stage = e->snap + e->snap_off;
    e->snap_off += (n + 1) & ~(size_t)1;
    if (cols != aos_stride && !d_extra) cudaMemsetAsync(stage, 0, n * sizeof(double), e->stream);
  }
  if (cols != aos_stride) {
    // partial rows (2D tensors: 4 in-plane + slot 4): assemble the whole row on the device
    if (d_extra) k_soa_to_aos<<<nblk((size_t)e->np, 256), 256, 0, e->stream>>>(d_extra, stage, rowmap, e->np, e->P.ld, 1, aos_stride, cols);
  }
  k_soa_to_aos<<<nblk((size_t)e->np * cols, 256), 256, 0, e->stream>>>(d, stage, rowmap, e->np, e->P.ld, cols, aos_stride, col0);
  if (stage != e->stage) {
    e->dl_copies.push_back({h, stage, n * sizeof(double)});
    return 0;
  }
  if (rows == 1) {
    e->h_rows.resize(n);
    CUDA_OK(cudaMemcpyAsync(e->h_rows.data(), e->stage, n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    for (int p = 0; p < e->np; p++) memcpy(h + (size_t)e->h_ids[p] * aos_stride, &e->h_rows[(size_t)p * aos_stride], sizeof(double) * aos_stride);
    return 0;
  }
  CUDA_OK(cudaMemcpyAsync(h, e->stage, n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CUDA_OK(cudaStreamSynchronize(e->stream));
  return 0;
}
static int get_ints(nlps_engine* e, int* h, const int* d, int rows) {
  if (!h || e->np == 0) return 0;
  if (e->dl_async && rows != 1) {
    int* stage = (int*)(e->snap + e->snap_off);
    e->snap_off += ((size_t)e->np / 2 + 2) & ~(size_t)1;
    k_unpermute_int<<<nblk(e->np, 256), 256, 0, e->stream>>>(d, stage, rows == 0 ? e->P.orig : nullptr, e->np);
    e->dl_copies.push_back({h, stage, sizeof(int) * e->np});
    return 0;
  }
  k_unpermute_int<<<nblk(e->np, 256), 256, 0, e->stream>>>(d, (int*)e->stage, rows == 0 ? e->P.orig : nullptr, e->np);
  if (rows == 1) {
    std::vector<int> tmp(e->np);
    CUDA_OK(cudaMemcpyAsync(tmp.data(), e->stage, sizeof(int) * e->np, cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    for (int p = 0; p < e->np; p++) h[e->h_ids[p]] = tmp[p];
    return 0;
  }
  CUDA_OK(cudaMemcpyAsync(h, e->stage, sizeof(int) * e->np, cudaMemcpyDeviceToHost, e->stream));
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
  if (e->host_fail && e->h_err[0] == 0) return 1;  // last_code was set where the failure happened
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

// physical re-sort of every particle array into the order given by plist; afterwards plist is the
// identity.  The results do not depend on when (or whether) this runs: cells are summed in the order of
// the caller's particle ids and everything else is per particle.
static void reorder_particles(nlps_engine* e) {
  const int np = e->np, ld = e->P.ld, D = e->D, T = e->T, DD = D * D;
  PartDev& P = e->P;
  const int* pl = e->G.plist;
  if (np == 0) { e->steps_since_sort = 0; return; }
  if (e->profile) cudaEventRecord(e->ev0, e->stream);
  auto gd = [&](double* f, int cols) {
    if (!f) return;
    k_gather_rows<double><<<nblk((size_t)np * cols, 256), 256, 0, e->stream>>>(f, e->stage, pl, np, ld, cols);
    cudaMemcpy2DAsync(f, sizeof(double) * ld, e->stage, sizeof(double) * ld, sizeof(double) * np, cols, cudaMemcpyDeviceToDevice, e->stream);
    e->launches++;
  };
  auto gi = [&](int* f, int cols) {
    k_gather_rows<int><<<nblk((size_t)np * cols, 256), 256, 0, e->stream>>>(f, (int*)e->stage, pl, np, ld, cols);
    cudaMemcpy2DAsync(f, sizeof(int) * ld, e->stage, sizeof(int) * ld, sizeof(int) * np, cols, cudaMemcpyDeviceToDevice, e->stream);
    e->launches++;
  };
  gd(P.x, D); gd(P.dis, D); gd(P.ddis, D); gd(P.vel, D); gd(P.acc, D); gd(P.lam, D);
  gd(P.beta, 1); gd(P.mass, 1); gd(P.vol0, 1); gd(P.rho, 1); gd(P.W, 1);
  gd(P.J_n, 1); gd(P.J_n1, 1); gd(P.eps_n, 1); gd(P.eps_n1, 1); gd(P.kap_n, 1); gd(P.kap_n1, 1);
  gd(P.F_n, DD); gd(P.F_n1, DD); gd(P.DF, DD); gd(P.be_n, T); gd(P.be_n1, T); gd(P.stress, T); gd(P.cep, DD);
  gd(P.Fs4, 1); gd(P.DFs4, 1); gd(P.area0, 1); gd(P.sstar, 1); gd(P.back, 3); gd(e->alme.bten, 4); gd(e->alme.cten, 4);
  gi(P.I0, 1); gi(P.nnodes, 1); gi(P.matidx, 1); gi(P.orig, 1); gi((int*)P.mask, e->W);
  k_after_sort<<<nblk(np, 256), 256, 0, e->stream>>>(P, e->G);
  e->launches++;
  if (e->profile) {
    cudaEventRecord(e->ev1, e->stream);
    cudaEventSynchronize(e->ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev0, e->ev1);
    e->k_ms[K_REORDER] += ms;
    e->k_n[K_REORDER]++;
  }
  e->steps_since_sort = 0;
  e->n_reorders++;
}

static SlabDev slab_dev(const nlps_engine* e) {
  SlabDev sl{};
  sl.on = e->slab_on;
  sl.axis = e->axis;
  const double roam = (e->band_cells - 3.5) * e->solver_dx;
  sl.own_lo = e->cut_lo; sl.own_hi = e->cut_hi;
  sl.lo = e->cut_lo - roam; sl.hi = e->cut_hi + roam;
  return sl;
}

// ---------------------------------------------------------------------------
// Slab halo exchange over the nodes within the band of each cut.  which: 0 occupancy, 1 M + momentum, 2 forces.
template <int D>
static int halo_exchange(nlps_engine* e, int which) {
  if (!e->slab_on) return 0;
  const int per = (which == 0) ? 1 : ((which == 1) ? 1 + D : D);
  nlps_msg msgs[2];
  int nm = 0;
  if (e->profile) cudaEventRecord(e->ev0, e->stream);
  if (e->p2p_on) {
    HaloP2P h{};
    const unsigned long long seq = ++e->halo_seq[which];
    int tot = 0;
    for (int s_ = 0; s_ < 2; s_++) {
      auto& sd = e->side[s_];
      const bool on = sd.peer >= 0 && sd.n > 0;
      h.ids[s_] = sd.ids;
      h.n[s_] = on ? sd.n : 0;
      // two receive buffers per exchange kind, taken in turn: a neighbour that is one exchange of the same kind ahead
      // (initialize_lme's search followed by the first step's, a stage called twice) writes into the other one
      const size_t off = ((size_t)(seq & 1ull) * 3 + which) * sd.n * (1 + D);
      h.peer_rbuf[s_] = on ? sd.peer_rbuf + off : nullptr;
      h.peer_flag[s_] = on ? sd.peer_flag + which : nullptr;
      h.my_rbuf[s_] = on ? sd.p2p_rbuf + off : nullptr;
      h.my_flag[s_] = on ? sd.p2p_flag + which : nullptr;
      tot += h.n[s_];
    }
    if (tot > 0) {
      k_halo_push<D><<<nblk(tot, 256), 256, 0, e->stream>>>(e->G, h, which, seq, e->p2p_done);
      // a neighbour may be busy for a while (set-up, a migration with a host sync): 30 s by default, NLPS_HALO_TIMEOUT_S
      static const long long tmo = (long long)((getenv("NLPS_HALO_TIMEOUT_S") ? atof(getenv("NLPS_HALO_TIMEOUT_S")) : 30.0) * 1.9e9);
      k_halo_pull<D><<<nblk(tot, 256), 256, 0, e->stream>>>(e->G, h, which, seq, e->err, tmo);
      e->launches += 2;
    }
    if (e->profile) {
      cudaEventRecord(e->ev1, e->stream);
      cudaEventSynchronize(e->ev1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e->ev0, e->ev1);
      e->k_ms[K_HALO] += ms;
      e->k_n[K_HALO]++;
    }
    return 0;
  }
  for (int s_ = 0; s_ < 2; s_++) {
    auto& h = e->side[s_];
    if (h.peer < 0 || h.n == 0) continue;
    k_halo_pack<D><<<nblk(h.n, 256), 256, 0, e->stream>>>(e->G, h.ids, h.n, which, h.sbuf);
    e->launches++;
    const unsigned long long bytes = sizeof(double) * (unsigned long long)h.n * per;
    msgs[nm++] = nlps_msg{h.peer, h.sbuf, bytes, h.rbuf, bytes};
  }
  int rc = comm_exchange(e->comm, nm, msgs, e->stream);
  for (int s_ = 0; s_ < 2 && !rc; s_++) {
    auto& h = e->side[s_];
    if (h.peer < 0 || h.n == 0) continue;
    k_halo_add<D><<<nblk(h.n, 256), 256, 0, e->stream>>>(e->G, h.ids, h.n, which, h.rbuf);
    e->launches++;
  }
  if (e->profile) {
    cudaEventRecord(e->ev1, e->stream);
    cudaEventSynchronize(e->ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev0, e->ev1);
    e->k_ms[K_HALO] += ms;
    e->k_n[K_HALO]++;
  }
  if (rc) { e->last_code = NLPS_ERR_CUDA; e->host_fail = 1; }
  return rc;
}

// Migration: mark, permute (stayers first), ship the tails to the neighbour slabs, append what arrives.
template <int D>
static int migrate_t(nlps_engine* e) {
  if (!e->slab_on) return 0;
  PartDev& P = e->P;
  const int np = e->np, ld = P.ld, T = e->T, DD = D * D;
  const SlabDev sl = slab_dev(e);
  cudaMemsetAsync(e->mig_cnt, 0, sizeof(int) * 16, e->stream);
  if (np) k_mig_mark<D><<<nblk(np, 256), 256, 0, e->stream>>>(e->mesh, P, sl, e->mig_dest, e->mig_cnt);
  // counts to the neighbours (slot 6: to lower, 7: to upper; 8: from lower, 9: from upper)
  cudaMemcpyAsync(e->mig_cnt + 6, e->mig_cnt + 1, sizeof(int) * 2, cudaMemcpyDeviceToDevice, e->stream);
  nlps_msg msgs[2];
  int nm = 0;
  for (int s_ = 0; s_ < 2; s_++)
    if (e->side[s_].peer >= 0) msgs[nm++] = nlps_msg{e->side[s_].peer, e->mig_cnt + 6 + s_, sizeof(int), e->mig_cnt + 8 + s_, sizeof(int)};
  if (comm_exchange(e->comm, nm, msgs, e->stream)) { e->last_code = NLPS_ERR_CUDA; return 1; }
  CUDA_OK(cudaMemcpyAsync(e->h_mig, e->mig_cnt, sizeof(int) * 16, cudaMemcpyDeviceToHost, e->stream));
  CUDA_OK(cudaStreamSynchronize(e->stream));
  const int n_stay = e->h_mig[0], n_out[2] = {e->h_mig[1], e->h_mig[2]}, n_in[2] = {e->h_mig[8], e->h_mig[9]};
  e->steps_since_migration = 0;
  // The decision to go on is taken by ALL slabs together: a slab that bailed out on its own would leave its
  // neighbours waiting in the row exchange below (or pairing that receive with a later halo message).
  int my_code = 0;
  if ((n_out[0] && e->side[0].peer < 0) || (n_out[1] && e->side[1].peer < 0)) {
    fprintf(stderr, "nlps_b200: slab %d: particles left the outer end of the slab range\n", e->rank);
    my_code = NLPS_ERR_SLAB_EXCURSION;
  }
  const int np_new = n_stay + n_in[0] + n_in[1];
  if (!my_code && (np_new > ld || std::max(std::max(n_out[0], n_out[1]), std::max(n_in[0], n_in[1])) > e->mig_cap)) {
    fprintf(stderr, "nlps_b200: slab %d: migration exceeds the capacity (%d rows, %d per message)\n", e->rank, ld, e->mig_cap);
    my_code = NLPS_ERR_SLAB_CAPACITY;
  }
  int all_ok = 1;
  if (comm_all_ok(e->comm, e->rank, e->world, my_code == 0, e->mig_cnt + 12, e->stream, &all_ok)) {
    e->last_code = NLPS_ERR_CUDA;
    e->host_fail = 1;
    return 1;
  }
  if (!all_ok) {  // every slab returns here, none enters the row exchange
    e->last_code = my_code ? my_code : NLPS_ERR_SLAB_CAPACITY;
    e->host_fail = 1;
    if (!my_code) fprintf(stderr, "nlps_b200: slab %d: another slab could not migrate its particles\n", e->rank);
    return 1;
  }
  const bool any = n_out[0] + n_out[1] + n_in[0] + n_in[1] > 0;
  if (n_out[0] + n_out[1]) {
    k_mig_perm<<<nblk(np, 256), 256, 0, e->stream>>>(np, e->mig_dest, e->mig_cnt, e->G.plist);
    reorder_particles(e);
    k_mig_inv<<<nblk(n_out[0] + n_out[1], 256), 256, 0, e->stream>>>(P, n_stay, n_out[0] + n_out[1], 0);
  }
  // column table of the travelling state (doubles first, then ints)
  std::vector<MigCol> tab;
  unsigned long long off = 0;
  auto addd = [&](double* f, int cols) { for (int c = 0; c < cols; c++) { tab.push_back(MigCol{f + (size_t)c * ld, 0, 0, off}); off += 8; } };
  auto addi = [&](int* f, int cols) { for (int c = 0; c < cols; c++) { tab.push_back(MigCol{f + (size_t)c * ld, 1, 0, off}); off += 4; } };
  addd(P.x, D); addd(P.dis, D); addd(P.ddis, D); addd(P.vel, D); addd(P.acc, D); addd(P.lam, D);
  addd(P.beta, 1); addd(P.mass, 1); addd(P.vol0, 1); addd(P.rho, 1); addd(P.W, 1);
  addd(P.J_n, 1); addd(P.J_n1, 1); addd(P.eps_n, 1); addd(P.eps_n1, 1); addd(P.kap_n, 1); addd(P.kap_n1, 1);
  addd(P.F_n, DD); addd(P.F_n1, DD); addd(P.DF, DD); addd(P.be_n, T); addd(P.be_n1, T); addd(P.stress, T); addd(P.cep, DD);
  addd(P.Fs4, 1); addd(P.DFs4, 1);
  if (P.area0) addd(P.area0, 1);
  if (P.back) addd(P.back, 3);
  addd(P.sstar, 1);
  addi(P.I0, 1); tab.back().is_int = 2;
  addi(P.nnodes, 1); addi(P.matidx, 1); addi(P.orig, 1); addi((int*)P.mask, e->W);
  const unsigned long long row_bytes = off;
  const int ncols = (int)tab.size();
  if (any) {
    CUDA_OK(cudaMemcpyAsync(e->mig_tab, tab.data(), sizeof(MigCol) * ncols, cudaMemcpyHostToDevice, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));  // tab is a host temporary
  }
  nm = 0;
  int row0 = n_stay;
  for (int s_ = 0; s_ < 2; s_++) {
    if (e->side[s_].peer < 0) continue;
    if (n_out[s_]) k_mig_copy<<<nblk((size_t)ncols * n_out[s_], 256), 256, 0, e->stream>>>(e->mig_tab, ncols, row0, n_out[s_], e->mig_sbuf[s_], 0, e->node_offset);
    row0 += n_out[s_];
    if (n_out[s_] || n_in[s_])
      msgs[nm++] = nlps_msg{e->side[s_].peer, e->mig_sbuf[s_], row_bytes * n_out[s_], e->mig_rbuf[s_], row_bytes * n_in[s_]};
  }
  if (comm_exchange(e->comm, nm, msgs, e->stream)) { e->last_code = NLPS_ERR_CUDA; return 1; }
  if (!any) return 0;
  row0 = n_stay;
  for (int s_ = 0; s_ < 2; s_++) {
    if (e->side[s_].peer < 0 || !n_in[s_]) continue;
    k_mig_copy<<<nblk((size_t)ncols * n_in[s_], 256), 256, 0, e->stream>>>(e->mig_tab, ncols, row0, n_in[s_], e->mig_rbuf[s_], 1, e->node_offset);
    row0 += n_in[s_];
  }
  e->np = np_new;
  P.np = np_new;
  if (n_in[0] + n_in[1]) k_mig_inv<<<nblk(n_in[0] + n_in[1], 256), 256, 0, e->stream>>>(P, n_stay, n_in[0] + n_in[1], 1);
  e->n_migrated_in += n_in[0] + n_in[1];
  e->steps_since_sort = 1 << 30;  // cell-sort the new population at the next search
  e->launches += 6;
  CUDA_OK(cudaStreamSynchronize(e->stream));
  return 0;
}

// launch record of the warp-per-cell kernels (nlps_cellwarp.cu)
static CwLaunch cw_launch_record(nlps_engine* e, const StepParams& sp) {
  CwLaunch L;
  L.m = e->mesh; L.P = e->P; L.G = e->G; L.sp = sp; L.cfg = e->cw; L.err = e->err; L.stream = e->stream;
  L.max_blocks = std::max(1, nblk((size_t)std::max(e->max_occ, 1), e->cw.warps));
  return L;
}
#define CW_TIMED(e, id, call)                                    \
  do {                                                           \
    if ((e)->profile) cudaEventRecord((e)->ev0, (e)->stream);    \
    if (call) (e)->last_code = NLPS_ERR_CUDA;                    \
    (e)->launches++;                                             \
    if ((e)->profile) {                                          \
      cudaEventRecord((e)->ev1, (e)->stream);                    \
      cudaEventSynchronize((e)->ev1);                            \
      float _ms = 0;                                             \
      cudaEventElapsedTime(&_ms, (e)->ev0, (e)->ev1);            \
      (e)->k_ms[id] += _ms;                                      \
      (e)->k_n[id]++;                                            \
    }                                                            \
  } while (0)

// proj != nullptr: implicit scheme, project that field instead of D_dis; lists_only_reuse: skip the whole search
// (second projection of the same step) and reuse lists, beta and lambda
template <int D>
static void stage_search_t(nlps_engine* e, int step, int update_I0, int do_predictor, const double* proj = nullptr,
                           int reuse = 0) {
  if (reuse) {
    StepParams sp = make_params(e, step, update_I0);
    sp.proj = proj;
    sp.reuse_lists = 1;
    if (e->kver == 2) {
      const CwLaunch L = cw_launch_record(e, sp);
      CW_TIMED(e, K_LME_P2G, cw_launch_lme_p2g(D, e->W, L, e->cws, e->sm_count, e->max_smem_optin, 0));
      return;
    }
    const int grid = std::max(1, nblk((size_t)e->max_occ, e->cfg.C));
#define CASE_WC(w, c) { auto kfn = k_lme_p2g<D, w, c>; LAUNCH_SMEM(e, K_LME_P2G, kfn, grid, e->cfg.threads, e->smemA, e->mesh, e->P, e->G, sp, e->cfg, e->err, 0, e->alme); }
#define CASE_W(w) case w: if (e->cache_pa) CASE_WC(w, true) else CASE_WC(w, false) break;
#define CASE_WF(w) case w: CASE_WC(w, false) break;
    if constexpr (D == 2) { switch (e->W) { CASE_W(1) CASE_W(2) } } else { switch (e->W) { CASE_WF(4) CASE_WF(8) } }
#undef CASE_WF
#undef CASE_W
#undef CASE_WC
    return;
  }
  if (e->slab_on && update_I0 && e->migrate_every > 0 && e->steps_since_migration >= e->migrate_every && !e->host_fail &&
      migrate_t<D>(e))
    e->host_fail = 1;
  e->steps_since_migration++;
  const int np = e->np, nn = e->nn;
  cudaMemsetAsync(e->G.cnt, 0, sizeof(int) * nn, e->stream);
  if (np) LAUNCH(e, K_SEARCH, k_search<D>, nblk(np, 256), 256, e->mesh, e->P, e->G, update_I0, slab_dev(e), e->err);
  halo_exchange<D>(e, 0);
  std::swap(e->G.dirty_cur, e->G.dirty_prev);
  LAUNCH(e, K_MARK, k_mark, nblk(nn, 256), 256, e->mesh, e->G);
  LAUNCH(e, K_MARK, k_live, nblk(nblk(nn, 256), 256), 256, e->G, nblk(nn, 256));
  LAUNCH(e, K_NODE_FLAGS, k_node_flags, nblk(nn, 256), 256, e->mesh, e->G);
  int nb = nblk(nn, SCAN_ITEMS);
  LAUNCH(e, K_SCAN1, k_scan_reduce, nb, 256, e->G.packed, e->G.scan_blk, nn, e->G.live);
  LAUNCH(e, K_SCAN2, k_scan_tops, 1, 1024, e->G.scan_blk, nb, e->G.n_active, e->G.n_occ, e->npart_check);
  LAUNCH(e, K_SCAN3, k_scan_apply, nb, 256, e->G.packed, e->G.scan_blk, e->G.cell_start, e->G.occ_pos, e->G.act_pos, nn, e->G.live);
  if (np) LAUNCH(e, K_FILL, k_cell_fill, nblk(np, 256), 256, e->P, e->G);
  LAUNCH(e, K_NODE_FINISH, k_node_finish, nblk(nn, 256), 256, e->mesh, e->P, e->G);
  if (e->reorder_every > 0 && e->steps_since_sort >= e->reorder_every) reorder_particles(e);
  e->steps_since_sort++;
  StepParams sp = make_params(e, step, update_I0);
  sp.proj = proj;
  if (e->kver == 2) {
    const CwLaunch L = cw_launch_record(e, sp);
    CW_TIMED(e, K_LME_P2G, cw_launch_lme_p2g(D, e->W, L, e->cws, e->sm_count, e->max_smem_optin, do_predictor));
    return;
  }
  const int grid = std::max(1, nblk((size_t)e->max_occ, e->cfg.C));
#define CASE_WC(w, c) { auto kfn = k_lme_p2g<D, w, c>; LAUNCH_SMEM(e, K_LME_P2G, kfn, grid, e->cfg.threads, e->smemA, e->mesh, e->P, e->G, sp, e->cfg, e->err, do_predictor, e->alme); }
#define CASE_W(w) case w: if (e->cache_pa) CASE_WC(w, true) else CASE_WC(w, false) break;
#define CASE_WF(w) case w: CASE_WC(w, false) break;
  // instantiated: 2D with 1-2 mask words (weights cached or not), 3D with 4-8 words (no weight cache)
  if constexpr (D == 2) { switch (e->W) { CASE_W(1) CASE_W(2) } } else { switch (e->W) { CASE_WF(4) CASE_WF(8) } }
#undef CASE_WF
#undef CASE_W
#undef CASE_WC
}
// node kernels: a thread per node with the slot-major partial sums, a warp per node with the cell-major ones (3D)
template <int D, int MODE>
static void launch_grid_disp(nlps_engine* e, int step, const int* ids0 = nullptr, int n0 = 0, const int* ids1 = nullptr, int n1 = 0) {
  const bool band = MODE == 2 && (ids0 || ids1);
  if (e->G.cm_sl && MODE != 2) {
    constexpr int LPN = (D == 3) ? 32 : 8;
    auto kf = k_grid_disp<D, MODE, LPN>;
    LAUNCH(e, K_GRID_DISP, kf, nblk((size_t)e->max_act * LPN, 128), 128, e->mesh, e->G, e->bc, step, ids0, n0, ids1, n1);
  } else {
    auto kf = k_grid_disp<D, MODE, 1>;
    LAUNCH(e, K_GRID_DISP, kf, nblk(band ? n0 + n1 : e->max_act, 128), 128, e->mesh, e->G, e->bc, step, ids0, n0, ids1, n1);
  }
}
template <int D, int MODE>
static void launch_grid_acc(nlps_engine* e, int step, const int* ids0 = nullptr, int n0 = 0, const int* ids1 = nullptr, int n1 = 0) {
  const bool band = MODE == 2 && (ids0 || ids1);
  if (e->G.cm_sl && MODE != 2) {
    constexpr int LPN = (D == 3) ? 32 : 8;
    auto kf = k_grid_acc<D, MODE, LPN>;
    LAUNCH(e, K_GRID_ACC, kf, nblk((size_t)e->max_act * LPN, 128), 128, e->mesh, e->G, e->grav, e->solver.num_steps, step, ids0, n0, ids1, n1);
  } else {
    auto kf = k_grid_acc<D, MODE, 1>;
    LAUNCH(e, K_GRID_ACC, kf, nblk(band ? n0 + n1 : e->max_act, 128), 128, e->mesh, e->G, e->grav, e->solver.num_steps, step, ids0, n0, ids1, n1);
  }
}

template <int D>
static void stage_p2g_mass_disp_t(nlps_engine* e, int step) {
  if (!e->slab_on) {
    launch_grid_disp<D, 0>(e, step);
  } else {
    // everything outside the halo bands is finished in the first pass; the band nodes after the exchange
    const int n0 = e->side[0].peer >= 0 ? e->side[0].n : 0, n1 = e->side[1].peer >= 0 ? e->side[1].n : 0;
    launch_grid_disp<D, 3>(e, step);
    halo_exchange<D>(e, 1);
    if (n0 + n1 > 0) launch_grid_disp<D, 2>(e, step, e->side[0].ids, n0, e->side[1].ids, n1);
  }
}
template <int D>
static void stage_kin_stress_t(nlps_engine* e, int step) {
  if (e->has_traction) {
    cudaMemsetAsync(e->P.trac, 0, sizeof(double) * (size_t)e->P.ld * D, e->stream);
    LAUNCH(e, K_TRACTION, k_traction<D>, nblk(e->neu.n_entries, 128), 128, e->P, e->neu, e->solver.thickness, step);
  }
  StepParams sp = make_params(e, step, 1);
  sp.implicit = e->implicit_on;
  e->n1_stale = 0;
  if (e->kver == 2) {
    const CwLaunch L = cw_launch_record(e, sp);
    if (e->uniform_mat == NLPS_MAT_NEO_HOOKEAN_WRIGGERS && !e->split_nh) {
      CW_TIMED(e, K_KIN_FORCE, cw_launch_kin(D, CW_KIN_FUSED, L, e->cws, e->sm_count, e->max_smem_optin, e->mat, e->has_traction));
    } else {  // gather (DF) -> stress (thread per particle) -> force sums
      CW_TIMED(e, K_KIN_GATHER, cw_launch_kin(D, CW_KIN_GATHER, L, e->cws, e->sm_count, e->max_smem_optin, e->mat, e->has_traction));
      CW_TIMED(e, K_STRESS, cw_launch_stress(D, L, e->mat, e->uniform_mat));
      CW_TIMED(e, K_KIN_FORCE, cw_launch_kin(D, CW_FORCE, L, e->cws, e->sm_count, e->max_smem_optin, e->mat, e->has_traction));
    }
    return;
  }
  const int grid = std::max(1, nblk((size_t)e->max_occ, e->cfg.C));
#define CASE_WMC(w, mt, c) { auto kfn = k_kin_force<D, w, mt, c>; LAUNCH_SMEM(e, K_KIN_FORCE, kfn, grid, e->cfg.threads, e->smemB, e->mesh, e->P, e->G, sp, e->cfg, e->err, e->has_traction, e->mat, e->alme); }
#define CASE_WM(w, mt) if (D == 2 && e->cache_pa) CASE_WMC(w, mt, (D == 2)) else CASE_WMC(w, mt, false)
#define CASE_W(w) case w: switch (e->uniform_mat) { case 0: CASE_WM(w, 0) break; case 1: CASE_WM(w, 1) break; case 2: CASE_WM(w, 2) break; default: CASE_WM(w, -1) break; } break;
  if constexpr (D == 2) { switch (e->W) { CASE_W(1) CASE_W(2) } } else { switch (e->W) { CASE_W(4) CASE_W(8) } }
#undef CASE_W
#undef CASE_WM
#undef CASE_WMC
}
template <int D>
static void stage_force_t(nlps_engine* e, int step) {
  if (!e->slab_on) {
    launch_grid_acc<D, 0>(e, step);
  } else {
    const int n0 = e->side[0].peer >= 0 ? e->side[0].n : 0, n1 = e->side[1].peer >= 0 ? e->side[1].n : 0;
    launch_grid_acc<D, 3>(e, step);
    halo_exchange<D>(e, 2);
    if (n0 + n1 > 0) launch_grid_acc<D, 2>(e, step, e->side[0].ids, n0, e->side[1].ids, n1);
  }
}
template <int D>
static void stage_g2p_t(nlps_engine* e, int step) {
  StepParams sp = make_params(e, step, 1);
  if (e->kver == 2) {
    const CwLaunch L = cw_launch_record(e, sp);
    CW_TIMED(e, K_G2P, cw_launch_g2p(D, L, e->cws, e->sm_count, e->max_smem_optin));
  } else {
  const int grid = std::max(1, nblk((size_t)e->max_occ, e->cfg.C));
#define CASE_W(w) case w: { auto kfn = k_g2p<D, w>; LAUNCH_SMEM(e, K_G2P, kfn, grid, e->cfg.threads, e->smemC, e->mesh, e->P, e->G, sp, e->cfg, e->alme); } break;
  if constexpr (D == 2) { switch (e->W) { CASE_W(1) CASE_W(2) } } else { switch (e->W) { CASE_W(4) CASE_W(8) } }
#undef CASE_W
  }
  if (!e->inert_synced) {
    if (e->np) k_sync_inert<D><<<nblk(e->np, 256), 256, 0, e->stream>>>(e->P, e->mat);
    e->inert_synced = 1;
  }
  // roll n+1 -> n (U-Verlet.c:1043-1081) as pointer swaps
  e->n1_stale = 1;
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

int nlps_b200_trim(int device) {
  cudaMemPool_t pool;
  if (cudaSetDevice(device) != cudaSuccess || cudaDeviceGetDefaultMemPool(&pool, device) != cudaSuccess) return 1;
  cudaDeviceSynchronize();
  return cudaMemPoolTrimTo(pool, 0) == cudaSuccess ? 0 : 1;
}

const char* nlps_b200_version(void) { return "nlps_b200 0.1 (sm_100a, fp64, explicit NPC-FS)"; }

// Process-wide cache of the peer-memory halo buffers (cudaMalloc + CUDA IPC export on my side, the opened mapping of
// the neighbour's buffer on the other): creating and tearing them down costs 0.15-0.45 s per engine (cudaMalloc,
// cudaIpcOpenMemHandle, cudaIpcCloseMemHandle, cudaFree all synchronise the device), far more than the halo traffic of a
// short run.  Entries live until the process exits (like the stream-ordered memory pool).  Every engine still does the
// handshake: a side reuses ITS buffer when it is large enough and says so through the generation number it sends; the
// neighbour keeps its mapping when the generation (and the exporter's pid) is the one it already opened.
struct P2PCacheEntry {
  int device, rank, peer, side;
  size_t bytes = 0;                      // of my receive buffer
  double* rbuf = nullptr;                // mine
  unsigned long long* flag = nullptr;
  cudaIpcMemHandle_t h_buf, h_flag;
  unsigned long long gen = 0;
  unsigned long long peer_gen = 0;       // generation of the neighbour's buffer that is mapped below (0 = none)
  long long peer_pid = 0;
  double* peer_rbuf = nullptr;
  unsigned long long* peer_flag = nullptr;
  bool in_use = false;                   // by a live engine (a second engine of the same slab side falls back to NCCL)
};
static std::vector<P2PCacheEntry*> g_p2p_cache;
static unsigned long long g_p2p_gen = 0;
static std::mutex g_p2p_mutex;
static P2PCacheEntry* p2p_cache_get(int device, int rank, int peer, int side) {
  std::lock_guard<std::mutex> lk(g_p2p_mutex);
  for (auto* c : g_p2p_cache)
    if (c->device == device && c->rank == rank && c->peer == peer && c->side == side) return c;
  auto* c = new P2PCacheEntry();
  c->device = device; c->rank = rank; c->peer = peer; c->side = side;
  g_p2p_cache.push_back(c);
  return c;
}

void nlps_b200_destroy(nlps_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  implicit_free(e);
  for (int s_ = 0; s_ < 2; s_++) {
    auto& h = e->side[s_];
    // peer-memory halo buffers and mappings belong to the process-wide cache (P2PCacheEntry): nothing to release here
    h.peer_rbuf = nullptr; h.peer_flag = nullptr; h.p2p_rbuf = nullptr; h.p2p_flag = nullptr;
    if (e->p2p_ce[s_]) {
      std::lock_guard<std::mutex> lk(g_p2p_mutex);
      ((P2PCacheEntry*)e->p2p_ce[s_])->in_use = false;
      e->p2p_ce[s_] = nullptr;
    }
  }
  for (void* p : e->allocs) pool_free(p, e->stream);
  if (e->stream) cudaStreamSynchronize(e->stream);
  if (e->h_err) cudaFreeHost(e->h_err);
  if (e->h_mig) cudaFreeHost(e->h_mig);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->dl_stream) { cudaStreamSynchronize(e->dl_stream); cudaStreamDestroy(e->dl_stream); }
  if (e->dl_ready) cudaEventDestroy(e->dl_ready);
  if (e->dl_done) cudaEventDestroy(e->dl_done);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

static int set_err(char* err, int len, const char* msg) {
  if (err && len > 0) { strncpy(err, msg, len - 1); err[len - 1] = 0; }
  fprintf(stderr, "nlps_b200_create: %s\n", msg);
  return 1;
}

static int upload_impl(nlps_engine* e, const nlps_particles* in, int rows);
static int create_impl(nlps_engine* e, const nlps_mesh* mesh, const nlps_solver* solver, int n_bounds,
                       const nlps_load* bounds, int n_neumann, const nlps_load* neumann, const double* gravity,
                       int n_materials, const nlps_material* materials, const nlps_particles* st,
                       const nlps_slab* slab, char* err, int err_len) {
  const int D = mesh->ndim, nn = mesh->n_nodes;
  const bool timing = getenv("NLPS_TIMING") != nullptr;
  auto now = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
  double t_mark = now();
  auto mark = [&](const char* what) {
    if (timing) { cudaStreamSynchronize(e->stream); fprintf(stderr, "  create: %-28s %.3f s\n", what, now() - t_mark); t_mark = now(); }
  };
  if (D != 2 && D != 3) return set_err(err, err_len, "ndim must be 2 or 3");
  if (!st->x_GC || !st->mass || !st->Vol_0 || !st->rho || !st->I0 || !st->MatIdx)
    return set_err(err, err_len, "x_GC, mass, Vol_0, rho, I0 and MatIdx are mandatory");
  for (int p = 0; p < st->n; p++)
    if (st->I0[p] < 0 || st->I0[p] >= nn) return set_err(err, err_len, "I0 out of range");
  // ---- which rows of `state` this engine holds
  std::vector<int> rows;  // state rows held (slab engines only)
  int np = st->n, ld = st->n;
  e->n_global = st->n;
  e->solver_dx = mesh->delta_x;
  if (slab) {
    if (slab->world < 1 || slab->rank < 0 || slab->rank >= slab->world || slab->axis < 0 || slab->axis >= D)
      return set_err(err, err_len, "bad slab description");
    if (slab->world > 1 && (!slab->cuts || !slab->comm)) return set_err(err, err_len, "slabs need cuts and a communicator");
    e->slab_on = 1;
    e->rank = slab->rank; e->world = slab->world; e->axis = slab->axis; e->comm = slab->comm;
    e->band_cells = slab->band_cells > 0 ? slab->band_cells : 6;
    if (e->band_cells < 4) return set_err(err, err_len, "band_cells must be >= 4");
    e->migrate_every = slab->migrate_every > 0 ? slab->migrate_every : 10;
    e->node_offset = slab->node_id_offset;
    e->n_global = slab->global_id ? slab->n_global : st->n;
    if (e->n_global < st->n && !slab->global_id) return set_err(err, err_len, "n_global smaller than the state");
    e->cut_lo = slab->rank > 0 ? slab->cuts[slab->rank - 1] : -1e300;
    e->cut_hi = slab->rank < slab->world - 1 ? slab->cuts[slab->rank] : 1e300;
    if (slab->rank > 0 && slab->rank < slab->world - 1 && !(e->cut_hi - e->cut_lo > 2.0 * e->band_cells * mesh->delta_x))
      return set_err(err, err_len, "slab narrower than two halo bands");
    for (int p = 0; p < st->n; p++) {
      if (slab->global_id && (slab->global_id[p] < 0 || slab->global_id[p] >= e->n_global))
        return set_err(err, err_len, "global particle id out of range");
      if (nlps_b200_slab_owner(mesh, slab->axis, slab->world, slab->cuts, st->I0[p]) == slab->rank) rows.push_back(p);
    }
    np = (int)rows.size();
    const double capf = slab->capacity_factor > 1.0 ? slab->capacity_factor : 1.3;
    // (an explicit capacity_factor is taken literally; the default leaves 1024 rows of slack for tiny clouds)
    ld = (int)(capf * std::max<double>(np, (double)e->n_global / slab->world)) + (slab->capacity_factor > 1.0 ? 1 : 1024);
  }
  e->D = D; e->T = (D == 2) ? 5 : 9; e->TB = e->T; e->nn = nn; e->np = np;
  e->solver = *solver;
  if (e->solver.quirk_transposed_eigvec < 0) e->solver.quirk_transposed_eigvec = (D == 2) ? 1 : 0;
  if (n_materials < 1 || n_materials > MAX_MATERIALS) return set_err(err, err_len, "1..8 materials supported");
  pool_setup(e->device);
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
  if (D == 3 && e->W < 4) e->W = 4;
  if (e->W > MAX_MASK_WORDS || (D == 2 && e->W > 2))
    return set_err(err, err_len, "2-ring larger than 64 (2D) / 256 (3D) nodes is not supported");
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
  mark("mesh upload");
  int maxr2t = 0, maxr1 = 0, maxr1t = 0;
  for (int i = 0; i < nn; i++) maxr1 = std::max(maxr1, mesh->ring1_ptr[i + 1] - mesh->ring1_ptr[i]);
  unsigned char* dq = nullptr;
  if (getenv("NLPS_HOST_TRANSPOSE")) {  // reference implementation of the same thing, kept for cross-checking
    std::vector<int> tp, ti;
    transpose_csr(nn, mesh->ring1_ptr, mesh->ring1_idx, tp, ti);
    if (dev_upload(e, &t1p, tp.data(), tp.size())) return 1;
    if (dev_upload(e, &t1i, ti.data(), ti.size())) return 1;
    CUDA_OK(cudaStreamSynchronize(e->stream));
    std::vector<unsigned char> qpos;
    transpose_csr(nn, mesh->ring2_ptr, mesh->ring2_idx, tp, ti, &qpos);
    for (int i = 0; i < nn; i++) maxr2t = std::max(maxr2t, tp[i + 1] - tp[i]);
    if (dev_upload(e, &t2p, tp.data(), tp.size())) return 1;
    if (dev_upload(e, &t2i, ti.data(), ti.size())) return 1;
    if (dev_upload(e, &dq, qpos.data(), qpos.size())) return 1;
    CUDA_OK(cudaStreamSynchronize(e->stream));
  } else {
    if (device_transpose(e, nn, r1p, r1i, mesh->ring1_ptr[nn], &t1p, &t1i, nullptr, &maxr1t)) return 1;
    if (device_transpose(e, nn, r2p, r2i, mesh->ring2_ptr[nn], &t2p, &t2i, &dq, &maxr2t)) return 1;
  }
  if (maxr2t > 255) return set_err(err, err_len, "transposed 2-ring larger than 255 nodes is not supported");
  mark("transposed adjacency");
  double* dsst = nullptr;
  if (dev_alloc(e, &dsst, (size_t)nn)) return 1;
  if (nn) k_sstar_nodes<<<nblk(nn, 256), 256, 0, e->stream>>>(dh, nn, solver->gamma_lme, e->neg_log_tol, dsst);
  e->mesh = MeshDev{nn, dX, r1p, r1i, r2p, r2i, t1p, t1i, t2p, t2i, dq, dh, dsst};
  {  // inverse of r2q: the slot a node holds in the 2-ring of each cell that lists it (cell-major partial sums)
    unsigned char *r2ts = nullptr, *r2pos = nullptr;
    if (dev_alloc(e, &r2ts, (size_t)std::max(mesh->ring2_ptr[nn], 1)) || dev_alloc(e, &r2pos, (size_t)std::max(mesh->ring2_ptr[nn], 1)))
      return 1;
    k_build_r2ts<<<nblk(nn, 256), 256, 0, e->stream>>>(nn, r2p, r2i, dq, t2p, r2ts, r2pos, D == 2 ? 1 : 0);
    e->mesh.r2ts = r2ts;
    e->mesh.r2pos = r2pos;
  }
  mark("transposed adjacency upload");
  e->max_occ = (int)std::min<long long>(nn, std::max(ld, 1));
  e->max_act = (int)std::min<long long>(nn, (long long)std::max(ld, 1) * maxr1);
  // ---- grid work arrays
  GridDev& G = e->G;
  if (dev_alloc(e, &G.M, nn) || dev_alloc(e, &G.UA, (size_t)nn * 2 * (D == 2 ? 2 : 4)) || dev_alloc(e, &G.F, (size_t)nn * D) ||
      dev_alloc(e, &G.active, nn) || dev_alloc(e, &G.fixed, nn) ||
      dev_alloc(e, &G.cnt, nn) || dev_alloc(e, &G.cursor, nn) || dev_alloc(e, &G.cell_start, nn) ||
      dev_alloc(e, &G.MOM, (size_t)nn * D) || dev_alloc(e, &G.rocc, nn) ||
      dev_alloc(e, &G.occ_blk, (size_t)nblk(nn, 256) + 1) || dev_alloc(e, &G.dirty_cur, (size_t)nblk(nn, 256) + 1) ||
      dev_alloc(e, &G.dirty_prev, (size_t)nblk(nn, 256) + 1) || dev_alloc(e, &G.live, (size_t)nblk(nn, 256) + 1) ||
      dev_alloc(e, &G.plist, ld) || dev_alloc(e, &G.act_list, nn) || dev_alloc(e, &G.n_active, 1) ||
      dev_alloc(e, &G.packed, nn) || dev_alloc(e, &G.scan_blk, (size_t)nblk(nn, SCAN_ITEMS) + 1) ||
      dev_alloc(e, &G.act_pos, nn) || dev_alloc(e, &G.occ_pos, nn) || dev_alloc(e, &G.occ_list, e->max_occ) ||
      dev_alloc(e, &G.n_occ, 1) || dev_alloc(e, &e->npart_check, 1) || dev_alloc(e, &e->err, 2) ||
      dev_alloc(e, &G.arank, nn) || dev_alloc(e, &G.occ_meta, (size_t)e->max_occ + 1))
    return 1;
  G.cap = maxr2t;
  G.max_act = e->max_act;
  G.w2t = (maxr2t + 31) / 32;
  // (sized for both layouts of the partial sums: slot-major max_act x cap x (1 + D), cell-major max_occ x maxr2 x 4)
  if (dev_alloc(e, &G.part, std::max((size_t)e->max_act * G.cap * (1 + D), (size_t)e->max_occ * maxr2 * 4))) return 1;
  if (dev_alloc(e, &G.occm, (size_t)e->max_act * G.w2t)) return 1;
  // ---- cell-block kernel configuration: cells per block, shared-memory budget
  {
    CUDA_OK(cudaDeviceGetAttribute(&e->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device));
    BlockCfg& c = e->cfg;
    c.SL = maxr2;
    c.threads = 128;
    if (const char* s_ = getenv("NLPS_THREADS")) c.threads = std::max(32, atoi(s_) / 32 * 32);
    e->cache_pa = (D == 2);
    if (const char* s_ = getenv("NLPS_CACHE_PA")) e->cache_pa = (D == 2) && atoi(s_) != 0;
    c.C = 32;
    if (const char* s_ = getenv("NLPS_CELLS_PER_BLOCK")) c.C = std::max(1, atoi(s_));
    // compact weight cache of the 3D cell phase of k_lme_p2g (NLPS_NCA=48 covers gamma = 6, n ~ 39-45; chunks holding
    // a particle with more neighbours recompute).  OFF by default: measured on the 64^3 cube it saves the exp of the
    // cell phase but its 25-33 KB of shared memory cost more in resident warps (4.31 -> 4.68 ms at 48, 5.63 at 64).
    c.NCA = c.NCB = 0;
    c.cellfast = 1;
    if (const char* s_ = getenv("NLPS_CELLFAST")) c.cellfast = atoi(s_);
    if (const char* s_ = getenv("NLPS_NCA")) c.NCA = (D == 3) ? std::max(0, atoi(s_)) : 0;
    auto sizes = [&](const BlockCfg& k, size_t& a, size_t& b, size_t& g) {
      if (D == 2) {
        switch (e->W) {
#define SZ_(d, w) case w: a = e->cache_pa ? LayoutA<d, w, true>(k).total : LayoutA<d, w, false>(k).total; \
                          b = e->cache_pa ? LayoutB<d, w, true>(k).total : LayoutB<d, w, false>(k).total; break;
          SZ_(2, 1) SZ_(2, 2) SZ_(2, 4) SZ_(2, 8)
        }
        g = LayoutC<2>(k).total;
      } else {
        switch (e->W) { SZ_(3, 1) SZ_(3, 2) SZ_(3, 4) SZ_(3, 8) }
#undef SZ_
        g = LayoutC<3>(k).total;
      }
    };
    // the group pipeline keeps the ring ids of the next group in NLPS_NID registers per thread and moves one metadata
    // record per thread
    c.threads = std::min(c.threads, 128);  // __launch_bounds__ of the cell-block kernels
    while (c.C > 1 && ((size_t)c.C * c.SL > (size_t)NLPS_NID * c.threads || c.C + 1 > c.threads)) c.C /= 2;
    if ((size_t)c.C * c.SL > (size_t)NLPS_NID * c.threads)
      return set_err(err, err_len, "2-ring too large for the cell-group pipeline");
    // keep at least two blocks per SM resident: halve the cells per block until the largest layout fits
    const size_t budget = std::min<size_t>((size_t)e->max_smem_optin, 100 * 1024);
    for (;;) {
      const double ppc = std::max(1.0, (double)std::max(np, 1) / std::max(1, e->max_occ));  // particles per occupied cell (lower bound)
      c.PCAP = std::max(32, (int)(c.C * std::max(ppc, (D == 2) ? 4.0 : 8.0) * 1.25 + 0.5));
      if (const char* s_ = getenv("NLPS_PCAP")) c.PCAP = std::max(1, atoi(s_));
      sizes(c, e->smemA, e->smemB, e->smemC);
      if (std::max(e->smemA, std::max(e->smemB, e->smemC)) <= budget || c.C == 1) break;
      c.C /= 2;
    }
    if (std::max(e->smemA, std::max(e->smemB, e->smemC)) > (size_t)e->max_smem_optin)
      return set_err(err, err_len, "2-ring too large for the shared-memory staging");
    c.magic = ((1u << 21) + c.SL - 1) / c.SL;
    for (unsigned q = 0; q < (unsigned)(c.C * c.SL); q++)
      if (((q * c.magic) >> 21) != q / c.SL) return set_err(err, err_len, "internal: pair-index division constant");
    // ---- warp-per-cell kernels (nlps_cellwarp.cu): three-dimensional decks
    {
      CwCfg& w = e->cw;
      w.SL = maxr2;
      w.warps = 4;
      w.NC = 48;  // gamma = 6: 33-48 neighbours; longer lists take two particle slots of the weight cache
      if (const char* s_ = getenv("NLPS_CW_NC")) w.NC = std::max(4, atoi(s_) & ~3);
      if (const char* s_ = getenv("NLPS_CW_WARPS")) w.warps = std::min(4, std::max(1, atoi(s_)));
      w.NC = std::max(w.NC, (((maxr2 + 3) & ~3) + 7) / 8);  // the longest possible list must fit the 8 slots of a chunk
      w.NC = (w.NC + 3) & ~3;
      w.CL = (maxr2 + 3) & ~3;
      w.W = e->W;
      e->kver = (D == 3) ? 2 : 1;
      if (const char* s_ = getenv("NLPS_KERNELS")) e->kver = (atoi(s_) == 1 || D != 3) ? 1 : 2;
      if (const char* s_ = getenv("NLPS_SPLIT_NH")) e->split_nh = atoi(s_) != 0;
      if (maxr2 > 256) e->kver = 1;  // slot ids of the compact lists are bytes
      // cell sums: cell-major with the warp-per-cell kernels (3D).  The block-per-cell-group kernels (2D) keep the slot-major
      // layout: their writers gain as much (lme -31 %, kin -19 % on BASELINE configs[1]) but a node there reads only 25
      // records and the gathering node kernels lose more (0.04 -> 0.18 ms each): 1.07 vs 1.18 ms per step,
      // profiles/ab/r02_ab_cellmajor_2d.txt.  NLPS_PART_CELLMAJOR=1 switches them over for A/B runs.
      e->G.cm_sl = (e->kver == 2) ? w.SL : 0;
      if (e->kver == 1 && getenv("NLPS_PART_CELLMAJOR") && atoi(getenv("NLPS_PART_CELLMAJOR")) != 0) e->G.cm_sl = w.SL;
      e->G.cm_w = w.W;
      if (e->kver == 2 && dev_alloc(e, &e->G.cum, (size_t)(e->max_occ + 1) * w.W)) return 1;
    }
    if (const char* s_ = getenv("NLPS_REORDER_EVERY")) e->reorder_every = atoi(s_);
    CUDA_OK(cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, e->device));
    if (const char* s_ = getenv("NLPS_GRID")) e->grid_override = std::max(1, atoi(s_));
  }
  mark("grid work arrays");
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
        if (neumann[b].ids[j] < 0 || neumann[b].ids[j] >= e->n_global) return set_err(err, err_len, "Neumann particle id out of range");
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
      if (s.type < 0 || s.type > NLPS_MAT_LADE_DUNCAN) return set_err(err, err_len, "unknown material type");
      hm[i] = MatParams{s.type, s.rho, s.E, s.nu, s.reference_pressure, s.kappa_0, s.hardening_modulus,
                        s.plastic_strain_0, s.phi_frictional, s.psi_frictional, s.exponent_hardening_ortiz,
                        s.cohesion, s.alpha_hardening_borja, s.a_hardening_borja[0], s.a_hardening_borja[1],
                        s.a_hardening_borja[2]};
      hm[i].voce_theta = s.theta_hardening_voce; hm[i].voce_K0 = s.k_0_hardening_voce;
      hm[i].voce_Kinf = s.k_inf_hardening_voce; hm[i].voce_delta = s.delta_hardening_voce;
      mat_hoist(hm[i], D);
      e->mat.m[i] = hm[i];
    }
    e->uniform_mat = hm[0].type;
    for (int i = 1; i < n_materials; i++)
      if (hm[i].type != hm[0].type) e->uniform_mat = -1;
  }
  mark("loads + materials");
  // ---- particles
  PartDev& P = e->P;
  P.np = np;
  P.ld = ld;
  const int T = e->T, DD = D * D, TBv = (D == 2) ? 5 : 9;
#define A_(f, c) if (dev_alloc(e, &P.f, (size_t)ld * (c))) return 1;
  A_(x, D) A_(dis, D) A_(ddis, D) A_(vel, D) A_(acc, D) A_(lam, D)
  A_(beta, 1) A_(mass, 1) A_(vol0, 1) A_(rho, 1) A_(W, 1)
  A_(J_n, 1) A_(J_n1, 1) A_(eps_n, 1) A_(eps_n1, 1) A_(kap_n, 1) A_(kap_n1, 1)
  A_(F_n, DD) A_(F_n1, DD) A_(DF, DD) A_(be_n, TBv) A_(be_n1, TBv) A_(stress, T) A_(cep, DD)
  A_(Fs4, 1) A_(DFs4, 1)
  A_(zi, 1) A_(ji, D * (D + 1) / 2) A_(gop, DD) A_(sstar, 1)
#undef A_
  if (dev_alloc(e, &P.clist, (size_t)ld * e->cw.CL)) return 1;
  P.trac = nullptr;
  P.area0 = nullptr;
  if (e->has_traction && dev_alloc(e, &P.trac, (size_t)ld * D)) return 1;
  if (e->has_traction && D == 3) {
    if (!st->Area_0)
      return set_err(err, err_len, "3D Neumann loads act on Particle.Phi.Area_0 (U-Verlet.c:847-849): state->Area_0 is NULL");
    if (dev_alloc(e, &P.area0, (size_t)ld)) return 1;
  }
  // GramsShapeFun (Type=aLME): metric and cut-off ellipsoid per particle (Nodes/aLME.c).  2D only, as the reference
  // (aLME.c:693-809 exits in 3D); the weights of the 2D kernels must stay cached in shared memory (the uncached cell
  // phases re-evaluate them with the scalar beta); single engine (the columns do not migrate between slabs)
  e->alme.bten = e->alme.cten = nullptr;
  if (e->solver.shape_function == NLPS_SHAPE_ALME) {
    if (D != 2) return set_err(err, err_len, "aLME shape functions are 2D only (Nodes/aLME.c:693-809 exits in 3D)");
    if (e->slab_on) return set_err(err, err_len, "aLME shape functions: single engine only (no slabs)");
    if (!e->cache_pa) return set_err(err, err_len, "aLME shape functions need NLPS_CACHE_PA=1 (the default in 2D)");
    if (dev_alloc(e, &e->alme.bten, (size_t)ld * 4) || dev_alloc(e, &e->alme.cten, (size_t)ld * 4)) return 1;
    CUDA_OK(cudaMemsetAsync(e->alme.bten, 0, sizeof(double) * (size_t)ld * 4, e->stream));
    CUDA_OK(cudaMemsetAsync(e->alme.cten, 0, sizeof(double) * (size_t)ld * 4, e->stream));
  } else if (e->solver.shape_function != NLPS_SHAPE_LME) {
    return set_err(err, err_len, "solver.shape_function must be NLPS_SHAPE_LME or NLPS_SHAPE_ALME");
  }
  P.back = nullptr;
  for (int i = 0; i < n_materials; i++)
    if (materials[i].type == NLPS_MAT_VON_MISES && !P.back) {
      if (dev_alloc(e, &P.back, (size_t)ld * 3)) return 1;
      CUDA_OK(cudaMemsetAsync(P.back, 0, sizeof(double) * (size_t)ld * 3, e->stream));
    }
  if (dev_alloc(e, &P.I0, ld) || dev_alloc(e, &P.nnodes, ld) || dev_alloc(e, &P.matidx, ld) ||
      dev_alloc(e, &P.orig, ld) || dev_alloc(e, &P.inv, std::max(e->n_global, 1)) || dev_alloc(e, &P.mask, (size_t)ld * e->W))
    return 1;
  e->stage_doubles = (size_t)std::max(ld, 1) * std::max(std::max(T, DD), (e->W + 1) / 2);
  if (dev_alloc(e, &e->stage, e->stage_doubles)) return 1;
  // defaults as allocate_U_vars__Fields__ leaves them (identity tensors, J = 1)
  auto fill = [&](double* a, size_t n, double v) { if (n) k_fill_d<<<nblk(n, 256), 256, 0, e->stream>>>(a, n, v); };
  for (int i = 0; i < D; i++) {
    fill(P.F_n + (size_t)(i * D + i) * ld, ld, 1.0);
    fill(P.F_n1 + (size_t)(i * D + i) * ld, ld, 1.0);
    fill(P.DF + (size_t)(i * D + i) * ld, ld, 1.0);
    fill(P.be_n + (size_t)(i * D + i) * ld, ld, 1.0);
    fill(P.be_n1 + (size_t)(i * D + i) * ld, ld, 1.0);
  }
  if (D == 2) { fill(P.be_n + (size_t)4 * ld, ld, 1.0); fill(P.be_n1 + (size_t)4 * ld, ld, 1.0); }
  fill(P.Fs4, ld, 1.0); fill(P.DFs4, ld, 1.0);
  fill(P.J_n, ld, 1.0); fill(P.J_n1, ld, 1.0);
  for (int p = 0; p < st->n; p++)
    if (st->MatIdx[p] < 0 || st->MatIdx[p] >= n_materials) return set_err(err, err_len, "MatIdx out of range");
  if (!e->slab_on) {
    k_iota<<<nblk(std::max(np, 1), 256), 256, 0, e->stream>>>(P.orig, P.inv, np);
    CUDA_OK(cudaStreamSynchronize(e->stream));
    mark("particle arrays");
    if (upload_impl(e, st, 0)) return 1;
    mark("particle upload");
    CUDA_OK(cudaMemcpyAsync(P.I0, st->I0, sizeof(int) * np, cudaMemcpyHostToDevice, e->stream));
    CUDA_OK(cudaMemcpyAsync(P.matidx, st->MatIdx, sizeof(int) * np, cudaMemcpyHostToDevice, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    return 0;
  }
  // ---- slab engine: gather the held rows on the host, ids = global ids
  {
    e->h_ids = rows;  // state-row indices: upload_impl(rows = 1) gathers by them
    // a caller that passes exactly its own particles (sub-mesh slabs: every row is held) needs no host gather:
    // the buffers go to the device as they are (pinned buffers then move at full link speed)
    bool all_rows = (int)rows.size() == st->n;
    for (int p = 0; all_rows && p < np; p++) all_rows = rows[p] == p;
    mark("slab particle arrays");
    // most rows held (a slab's own particles plus a margin): upload whole buffers, gather the held rows on the device
    int mode = all_rows ? 2 : 1;
    if (!all_rows && (size_t)st->n * std::max(e->T, D * D) <= e->stage_doubles) {
      if (dev_upload(e, &e->d_rows, rows.data(), rows.size())) return 1;
      e->up_rows = st->n;
      mode = 3;
    }
    if (upload_impl(e, st, mode)) return 1;
    mark(mode == 2 ? "slab field upload (direct)" : mode == 3 ? "slab field upload (device gather)" : "slab field upload (host gather)");
    std::vector<int> i0(std::max(np, 1)), mi(std::max(np, 1)), gid(std::max(np, 1));
    for (int p = 0; p < np; p++) {
      i0[p] = st->I0[rows[p]];
      mi[p] = st->MatIdx[rows[p]];
      gid[p] = slab->global_id ? slab->global_id[rows[p]] : rows[p];
    }
    CUDA_OK(cudaMemcpyAsync(P.I0, i0.data(), sizeof(int) * np, cudaMemcpyHostToDevice, e->stream));
    CUDA_OK(cudaMemcpyAsync(P.matidx, mi.data(), sizeof(int) * np, cudaMemcpyHostToDevice, e->stream));
    CUDA_OK(cudaMemcpyAsync(P.orig, gid.data(), sizeof(int) * np, cudaMemcpyHostToDevice, e->stream));
    k_fill_i<<<nblk(std::max(e->n_global, 1), 256), 256, 0, e->stream>>>(P.inv, (size_t)e->n_global, -1);
    if (np) k_set_inv<<<nblk(np, 256), 256, 0, e->stream>>>(P.orig, P.inv, np);
    CUDA_OK(cudaStreamSynchronize(e->stream));
    e->h_ids = gid;
    e->h_ids.resize(np);
  }
  mark("slab particle upload");
  // ---- halo node lists (identical on both sides of a cut) and exchange buffers
  for (int s_ = 0; s_ < 2; s_++) {
    const int peer = s_ == 0 ? e->rank - 1 : e->rank + 1;
    if (peer < 0 || peer >= e->world) continue;
    const double cut = s_ == 0 ? e->cut_lo : e->cut_hi;
    const int n = nlps_b200_slab_halo_nodes(mesh, e->axis, cut, e->band_cells, nullptr);
    std::vector<int> ids(std::max(n, 1));
    nlps_b200_slab_halo_nodes(mesh, e->axis, cut, e->band_cells, ids.data());
    auto& h = e->side[s_];
    h.peer = peer;
    h.n = n;
    if (dev_upload(e, &h.ids, ids.data(), (size_t)n) || dev_alloc(e, &h.sbuf, (size_t)n * (1 + D)) ||
        dev_alloc(e, &h.rbuf, (size_t)n * (1 + D)))
      return 1;
    CUDA_OK(cudaStreamSynchronize(e->stream));
    // band flags: the node kernels leave these nodes to the pass after the exchange
    if (!e->G.band) {
      if (dev_alloc(e, &e->G.band, (size_t)e->nn)) return 1;
      CUDA_OK(cudaMemsetAsync(e->G.band, 0, (size_t)e->nn, e->stream));
    }
    if (n > 0) k_set_flags<<<nblk(n, 256), 256, 0, e->stream>>>(h.ids, n, e->G.band);
  }
  if (e->slab_on && !e->G.band) {  // a slab without neighbours (world = 1)
    if (dev_alloc(e, &e->G.band, (size_t)e->nn)) return 1;
    CUDA_OK(cudaMemsetAsync(e->G.band, 0, (size_t)e->nn, e->stream));
  }
  // ---- migration buffers
  {
    const size_t row_bytes = 8 * (size_t)(6 * D + 11 + 4 * DD + 3 * T + 2 + 1 + (P.area0 ? 1 : 0) + (P.back ? 3 : 0)) + 4 * (size_t)(4 + e->W);
    e->mig_cap = std::max(4096, ld / 8);
    if (dev_alloc(e, &e->mig_dest, ld) || dev_alloc(e, &e->mig_cnt, 16) || dev_alloc(e, &e->mig_tab, 256)) return 1;
    for (int s_ = 0; s_ < 2; s_++)
      if (e->side[s_].peer >= 0 &&
          (dev_alloc(e, &e->mig_sbuf[s_], row_bytes * e->mig_cap) || dev_alloc(e, &e->mig_rbuf[s_], row_bytes * e->mig_cap)))
        return 1;
    CUDA_OK(cudaMallocHost(&e->h_mig, 16 * sizeof(int)));
    CUDA_OK(cudaStreamSynchronize(e->stream));
  }
  mark("halo + migration buffers");
  // ---- peer-memory halo path (NVLink stores through CUDA IPC mappings): needs the NCCL transport (the handles
  // travel over it) and one process per GPU; any failure leaves the NCCL path in place
  if (e->comm && e->comm->is_nccl && e->world > 1 && !(getenv("NLPS_P2P") && atoi(getenv("NLPS_P2P")) == 0)) {
    struct Blob { cudaIpcMemHandle_t buf, flag; int ok; int pad; unsigned long long gen; long long pid; };
    Blob mine[2], theirs[2];
    memset(mine, 0, sizeof(mine));
    memset(theirs, 0, sizeof(theirs));
    P2PCacheEntry* ce[2] = {nullptr, nullptr};
    bool ok = true;
    for (int s_ = 0; s_ < 2 && ok; s_++) {
      auto& h = e->side[s_];
      if (h.peer < 0) continue;
      const size_t nb = sizeof(double) * 2 * 3 * (size_t)std::max(h.n, 1) * (1 + D);
      P2PCacheEntry* c = ce[s_] = p2p_cache_get(e->device, e->rank, h.peer, s_);
      {
        std::lock_guard<std::mutex> lk(g_p2p_mutex);
        if (c->in_use) { ok = false; ce[s_] = nullptr; mine[s_].ok = 0; break; }
        c->in_use = true;
        e->p2p_ce[s_] = c;
      }
      if (!c->rbuf || c->bytes < nb) {
        // (a smaller buffer of an earlier engine stays allocated: the neighbour may still have it mapped)
        double* rb = nullptr;
        unsigned long long* fl = nullptr;
        ok = cudaMalloc(&rb, nb) == cudaSuccess && cudaMalloc(&fl, 4 * sizeof(unsigned long long)) == cudaSuccess &&
             cudaIpcGetMemHandle(&c->h_buf, rb) == cudaSuccess && cudaIpcGetMemHandle(&c->h_flag, fl) == cudaSuccess;
        if (ok) {
          std::lock_guard<std::mutex> lk(g_p2p_mutex);
          c->rbuf = rb; c->flag = fl; c->bytes = nb; c->gen = ++g_p2p_gen;
        }
      }
      if (ok) {
        // the neighbour's previous engine has pushed everything my previous engine waited for: safe to reset
        ok = cudaMemsetAsync(c->rbuf, 0, nb, e->stream) == cudaSuccess &&
             cudaMemsetAsync(c->flag, 0, 4 * sizeof(unsigned long long), e->stream) == cudaSuccess;
        h.p2p_rbuf = c->rbuf; h.p2p_flag = c->flag;
        mine[s_].buf = c->h_buf; mine[s_].flag = c->h_flag; mine[s_].gen = c->gen; mine[s_].pid = (long long)getpid();
      }
      mine[s_].ok = ok ? 1 : 0;
    }
    CUDA_OK(cudaStreamSynchronize(e->stream));  // buffers are zero before anybody learns (again) where they are
    // exchange the handles with the two neighbours over the transport (device bounce buffers)
    Blob *d_mine = nullptr, *d_theirs = nullptr;
    if (dev_alloc(e, &d_mine, 2) || dev_alloc(e, &d_theirs, 2)) return 1;
    CUDA_OK(cudaMemcpyAsync(d_mine, mine, sizeof(mine), cudaMemcpyHostToDevice, e->stream));
    nlps_msg msgs[2];
    int nm = 0;
    for (int s_ = 0; s_ < 2; s_++)
      if (e->side[s_].peer >= 0) msgs[nm++] = nlps_msg{e->side[s_].peer, d_mine + s_, sizeof(Blob), d_theirs + s_, sizeof(Blob)};
    if (comm_exchange(e->comm, nm, msgs, e->stream)) return 1;
    CUDA_OK(cudaMemcpyAsync(theirs, d_theirs, sizeof(theirs), cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    for (int s_ = 0; s_ < 2; s_++) {
      auto& h = e->side[s_];
      if (h.peer < 0) continue;
      if (!mine[s_].ok || !theirs[s_].ok || !ce[s_]) { ok = false; continue; }
      P2PCacheEntry* c = ce[s_];
      if (c->peer_rbuf && c->peer_gen == theirs[s_].gen && c->peer_pid == theirs[s_].pid) {
        h.peer_rbuf = c->peer_rbuf;  // the mapping opened by an earlier engine
        h.peer_flag = c->peer_flag;
        continue;
      }
      if (c->peer_rbuf) {  // the neighbour moved to a new buffer
        cudaIpcCloseMemHandle(c->peer_rbuf);
        cudaIpcCloseMemHandle(c->peer_flag);
        c->peer_rbuf = nullptr; c->peer_flag = nullptr; c->peer_gen = 0;
      }
      if (cudaIpcOpenMemHandle((void**)&h.peer_rbuf, theirs[s_].buf, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle((void**)&h.peer_flag, theirs[s_].flag, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        h.peer_rbuf = nullptr; h.peer_flag = nullptr;
        ok = false;
      } else {
        c->peer_rbuf = h.peer_rbuf; c->peer_flag = h.peer_flag; c->peer_gen = theirs[s_].gen; c->peer_pid = theirs[s_].pid;
      }
    }
    // everybody must take the same path: the verdict travels along the chain of slabs (world - 1 rounds of
    // neighbour exchanges of min(mine, theirs) reach every rank)
    int *d_v = nullptr, *d_g = nullptr;
    if (dev_alloc(e, &d_v, 2) || dev_alloc(e, &d_g, 2)) return 1;
    int verdict = ok ? 1 : 0;
    for (int round = 0; round < e->world - 1; round++) {
      int v2[2] = {verdict, verdict}, got[2] = {1, 1};
      CUDA_OK(cudaMemcpyAsync(d_v, v2, sizeof(v2), cudaMemcpyHostToDevice, e->stream));
      nm = 0;
      for (int s_ = 0; s_ < 2; s_++)
        if (e->side[s_].peer >= 0) msgs[nm++] = nlps_msg{e->side[s_].peer, d_v + s_, sizeof(int), d_g + s_, sizeof(int)};
      if (comm_exchange(e->comm, nm, msgs, e->stream)) return 1;
      CUDA_OK(cudaMemcpyAsync(got, d_g, sizeof(got), cudaMemcpyDeviceToHost, e->stream));
      CUDA_OK(cudaStreamSynchronize(e->stream));
      for (int s_ = 0; s_ < 2; s_++)
        if (e->side[s_].peer >= 0) verdict = std::min(verdict, got[s_]);
    }
    if (!verdict) {
      if (e->rank == 0) fprintf(stderr, "nlps_b200: peer-memory halo path unavailable, using NCCL send/recv\n");
      // (buffers and mappings are released by destroy)
    } else {
      if (dev_alloc(e, &e->p2p_done, 4)) return 1;
      e->p2p_on = 1;
      mark("peer-memory halo setup");
    }
  }
  // the first collective of a communicator pays NCCL's lazy channel set-up (hundreds of milliseconds): take the
  // "did every slab succeed" reduction of the migrations once here, not inside the first migration of a run
  if (e->slab_on && e->world > 1) {
    int all_ok = 1;
    if (comm_all_ok(e->comm, e->rank, e->world, 1, e->mig_cnt + 12, e->stream, &all_ok) || !all_ok)
      return set_err(err, err_len, "slab communicator: the collective of the migration check failed");
    mark("collective warm-up");
  }
  return 0;
}

static nlps_engine* create_any(const nlps_mesh* mesh, const nlps_solver* solver, int n_bounds, const nlps_load* bounds,
                               int n_neumann, const nlps_load* neumann, const double* gravity, int n_materials,
                               const nlps_material* materials, const nlps_particles* state, const nlps_slab* slab,
                               int device, char* err, int err_len) {
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
  if (create_impl(e, mesh, solver, n_bounds, bounds, n_neumann, neumann, gravity, n_materials, materials, state, slab,
                  err, err_len)) {
    nlps_b200_destroy(e);
    return nullptr;
  }
  return e;
}

nlps_engine* nlps_b200_create(const nlps_mesh* mesh, const nlps_solver* solver, int n_bounds, const nlps_load* bounds,
                              int n_neumann, const nlps_load* neumann, const double* gravity, int n_materials,
                              const nlps_material* materials, const nlps_particles* state, int device, char* err,
                              int err_len) {
  return create_any(mesh, solver, n_bounds, bounds, n_neumann, neumann, gravity, n_materials, materials, state, nullptr,
                    device, err, err_len);
}

nlps_engine* nlps_b200_create_slab(const nlps_mesh* mesh, const nlps_solver* solver, int n_bounds,
                                   const nlps_load* bounds, int n_neumann, const nlps_load* neumann,
                                   const double* gravity, int n_materials, const nlps_material* materials,
                                   const nlps_particles* state, const nlps_slab* slab, int device, char* err,
                                   int err_len) {
  if (!slab) {
    set_err(err, err_len, "nlps_b200_create_slab: slab description missing");
    return nullptr;
  }
  return create_any(mesh, solver, n_bounds, bounds, n_neumann, neumann, gravity, n_materials, materials, state, slab,
                    device, err, err_len);
}

static int upload_impl(nlps_engine* e, const nlps_particles* in, int rows) {
  if (in->b_e_n || in->b_e_n1 || in->EPS_n || in->EPS_n1 || in->Kappa_n || in->Kappa_n1) e->inert_synced = 0;
  const int D = e->D, T = e->T, DD = D * D;
  PartDev& P = e->P;
  if (put_field(e, in->x_GC, P.x, D, D, 0, rows) || put_field(e, in->dis, P.dis, D, D, 0, rows) ||
      put_field(e, in->D_dis, P.ddis, D, D, 0, rows) || put_field(e, in->vel, P.vel, D, D, 0, rows) ||
      put_field(e, in->acc, P.acc, D, D, 0, rows) || put_field(e, in->lambda, P.lam, D, D, 0, rows))
    return 1;
  if (put_field(e, in->F_n, P.F_n, DD, T, 0, rows) || put_field(e, in->F_n1, P.F_n1, DD, T, 0, rows) ||
      put_field(e, in->DF, P.DF, DD, T, 0, rows))
    return 1;
  if (D == 2) {
    if (put_field(e, in->F_n, P.Fs4, 1, T, 4, rows) || put_field(e, in->DF, P.DFs4, 1, T, 4, rows)) return 1;
  }
  if (put_field(e, in->b_e_n, P.be_n, T, T, 0, rows) || put_field(e, in->b_e_n1, P.be_n1, T, T, 0, rows) ||
      put_field(e, in->Stress, P.stress, T, T, 0, rows) || put_field(e, in->C_ep, P.cep, DD, DD, 0, rows))
    return 1;
  if (put_field(e, in->J_n, P.J_n, 1, 1, 0, rows) || put_field(e, in->J_n1, P.J_n1, 1, 1, 0, rows) ||
      put_field(e, in->mass, P.mass, 1, 1, 0, rows) || put_field(e, in->rho, P.rho, 1, 1, 0, rows) ||
      put_field(e, in->Vol_0, P.vol0, 1, 1, 0, rows) || put_field(e, in->W, P.W, 1, 1, 0, rows) ||
      put_field(e, in->EPS_n, P.eps_n, 1, 1, 0, rows) || put_field(e, in->EPS_n1, P.eps_n1, 1, 1, 0, rows) ||
      put_field(e, in->Kappa_n, P.kap_n, 1, 1, 0, rows) || put_field(e, in->Kappa_n1, P.kap_n1, 1, 1, 0, rows) ||
      (e->alme.bten ? put_field(e, in->Beta, e->alme.bten, 4, 4, 0, rows) : put_field(e, in->Beta, P.beta, 1, 1, 0, rows)) ||
      (e->alme.cten && put_field(e, in->Cut_off_Ellipsoid, e->alme.cten, 4, 4, 0, rows)) ||
      put_field(e, in->Area_0, P.area0, 1, 1, 0, rows) || put_field(e, in->Back_stress, P.back, 3, 3, 0, rows))
    return 1;
  if (e->np) k_sstar_particles<<<nblk(e->np, 256), 256, 0, e->stream>>>(P.beta, e->np, e->neg_log_tol, P.sstar);
  return 0;
}

int nlps_b200_upload(nlps_engine* e, const nlps_particles* in) {
  cudaSetDevice(e->device);
  if (!e->slab_on) return upload_impl(e, in, 0);
  if (refresh_ids(e)) return 1;  // rows of `in` are indexed by global id
  return upload_impl(e, in, 1);
}

static int download_impl(nlps_engine* e, nlps_particles* out, int rows) {
  const int D = e->D, T = e->T, DD = D * D;
  PartDev& P = e->P;
  if (get_field(e, out->x_GC, P.x, D, D, 0, nullptr, rows) || get_field(e, out->dis, P.dis, D, D, 0, nullptr, rows) ||
      get_field(e, out->D_dis, P.ddis, D, D, 0, nullptr, rows) || get_field(e, out->vel, P.vel, D, D, 0, nullptr, rows) ||
      get_field(e, out->acc, P.acc, D, D, 0, nullptr, rows) || get_field(e, out->lambda, P.lam, D, D, 0, nullptr, rows))
    return 1;
  // after a completed step the *_n1 host fields receive the rolled state, as the reference's copy roll leaves them
  const bool st_ = e->n1_stale != 0;
  const double *F1 = st_ ? P.F_n : P.F_n1, *J1 = st_ ? P.J_n : P.J_n1, *be1 = st_ ? P.be_n : P.be_n1;
  const double *eps1 = st_ ? P.eps_n : P.eps_n1, *kap1 = st_ ? P.kap_n : P.kap_n1;
  if (D == 2) {
    if (get_field(e, out->F_n, P.F_n, DD, T, 0, P.Fs4, rows) || get_field(e, out->F_n1, F1, DD, T, 0, P.Fs4, rows) ||
        get_field(e, out->DF, P.DF, DD, T, 0, P.DFs4, rows))
      return 1;
  } else {
    if (get_field(e, out->F_n, P.F_n, DD, T, 0, nullptr, rows) || get_field(e, out->F_n1, F1, DD, T, 0, nullptr, rows) ||
        get_field(e, out->DF, P.DF, DD, T, 0, nullptr, rows))
      return 1;
  }
  if (get_field(e, out->b_e_n, P.be_n, T, T, 0, nullptr, rows) || get_field(e, out->b_e_n1, be1, T, T, 0, nullptr, rows) ||
      get_field(e, out->Stress, P.stress, T, T, 0, nullptr, rows) || get_field(e, out->C_ep, P.cep, DD, DD, 0, nullptr, rows))
    return 1;
  if (get_field(e, out->J_n, P.J_n, 1, 1, 0, nullptr, rows) || get_field(e, out->J_n1, J1, 1, 1, 0, nullptr, rows) ||
      get_field(e, out->mass, P.mass, 1, 1, 0, nullptr, rows) || get_field(e, out->rho, P.rho, 1, 1, 0, nullptr, rows) ||
      get_field(e, out->Vol_0, P.vol0, 1, 1, 0, nullptr, rows) || get_field(e, out->W, P.W, 1, 1, 0, nullptr, rows) ||
      get_field(e, out->EPS_n, P.eps_n, 1, 1, 0, nullptr, rows) || get_field(e, out->EPS_n1, eps1, 1, 1, 0, nullptr, rows) ||
      get_field(e, out->Kappa_n, P.kap_n, 1, 1, 0, nullptr, rows) || get_field(e, out->Kappa_n1, kap1, 1, 1, 0, nullptr, rows) ||
      (e->alme.bten ? get_field(e, out->Beta, e->alme.bten, 4, 4, 0, nullptr, rows) : get_field(e, out->Beta, P.beta, 1, 1, 0, nullptr, rows)) ||
      (e->alme.cten && get_field(e, out->Cut_off_Ellipsoid, e->alme.cten, 4, 4, 0, nullptr, rows)) ||
      (P.back && get_field(e, out->Back_stress, P.back, 3, 3, 0, nullptr, rows)))
    return 1;
  if (get_ints(e, out->I0, P.I0, rows) || get_ints(e, out->NumberNodes, P.nnodes, rows)) return 1;
  return 0;
}

int nlps_b200_download(nlps_engine* e, nlps_particles* out) {
  cudaSetDevice(e->device);
  if (!e->slab_on) return download_impl(e, out, 0);
  if (out->n < e->n_global) {
    fprintf(stderr, "nlps_b200_download: slab engines write rows by global id: out->n must be n_global\n");
    return 1;
  }
  if (refresh_ids(e)) return 1;
  return download_impl(e, out, 1);
}

int nlps_b200_local_count(nlps_engine* e) { return e->np; }

const char* nlps_b200_transport(nlps_engine* e) {
  if (!e->slab_on || e->world <= 1) return "none (single slab)";
  if (e->p2p_on) return "peer-memory stores over NVLink (CUDA IPC mappings of the neighbours' halo buffers; NCCL for set-up and migration)";
  if (e->comm && e->comm->is_nccl) return "ncclSend/ncclRecv (grouped, on the engine's stream)";
  return "caller-supplied exchange function";
}

int nlps_b200_download_local(nlps_engine* e, nlps_particles* out, int* ids) {
  cudaSetDevice(e->device);
  if (out->n < e->np) return 1;
  if (refresh_ids(e)) return 1;
  if (ids) memcpy(ids, e->h_ids.data(), sizeof(int) * e->np);
  // compact rows in slot order (MatIdx is not part of download(): fetch it here for completeness of a row)
  if (out->MatIdx && e->np) {
    CUDA_OK(cudaMemcpyAsync(out->MatIdx, e->P.matidx, sizeof(int) * e->np, cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
  }
  return download_impl(e, out, 2);
}

int nlps_b200_migrate(nlps_engine* e) {
  cudaSetDevice(e->device);
  int rc = (e->D == 2) ? migrate_t<2>(e) : migrate_t<3>(e);
  if (rc) e->host_fail = 1;
  return rc ? 1 : poll_error(e);
}
long long nlps_b200_migrated_count(nlps_engine* e) { return e->n_migrated_in; }

int nlps_b200_comm_unique_id(char id[128]) {
  ncclUniqueId u;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  NcclApi* N = nccl_api();
  if (!N || N->GetUniqueId(&u) != ncclSuccess) return 1;
  memcpy(id, &u, 128);
  return 0;
}
nlps_comm* nlps_b200_comm_create_nccl(const char id[128], int rank, int world, int device) {
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  ncclUniqueId u;
  memcpy(&u, id, 128);
  nlps_comm* c = new nlps_comm();
  c->rank = rank; c->world = world; c->is_nccl = 1;
  NcclApi* N = nccl_api();
  ncclResult_t r = N ? N->CommInitRank(&c->nccl, world, u, rank) : ncclSystemError;
  if (r != ncclSuccess) {
    fprintf(stderr, "nlps_b200_comm_create_nccl: %s\n", N ? N->GetErrorString(r) : "NCCL library not available");
    delete c;
    return nullptr;
  }
  return c;
}
nlps_comm* nlps_b200_comm_create_custom(int rank, int world, nlps_exchange_fn fn, void* user) {
  if (!fn) return nullptr;
  nlps_comm* c = new nlps_comm();
  c->rank = rank; c->world = world; c->fn = fn; c->user = user;
  return c;
}
void nlps_b200_comm_destroy(nlps_comm* c) {
  if (!c) return;
  if (c->is_nccl && c->nccl) nccl_api()->CommDestroy(c->nccl);
  delete c;
}
int nlps_b200_memcpy_d2d(void* dst, const void* src, unsigned long long bytes, void* cuda_stream) {
  if (!bytes) return 0;
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)cuda_stream) == cudaSuccess ? 0 : 1;
}
int nlps_b200_stream_sync(void* cuda_stream) { return cudaStreamSynchronize((cudaStream_t)cuda_stream) == cudaSuccess ? 0 : 1; }

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
  if (e->np == 0) return 0;
  int* tmp = nullptr;
  size_t n = (size_t)e->np * cap;
  CUDA_OK(cudaMalloc(&tmp, n * sizeof(int)));
  k_expand_lists<<<nblk(e->np, 128), 128, 0, e->stream>>>(e->mesh, e->P, e->W, cap, tmp, e->slab_on);
  cudaError_t st = cudaSuccess;
  if (!e->slab_on) {
    st = cudaMemcpyAsync(lists, tmp, n * sizeof(int), cudaMemcpyDeviceToHost, e->stream);
    if (st == cudaSuccess) st = cudaStreamSynchronize(e->stream);
    if (st == cudaSuccess && counts && get_ints(e, counts, e->P.nnodes, 0)) st = cudaErrorUnknown;
  } else {  // rows of the caller's arrays are indexed by global id
    std::vector<int> h(n);
    st = cudaMemcpyAsync(h.data(), tmp, n * sizeof(int), cudaMemcpyDeviceToHost, e->stream);
    if (st == cudaSuccess) st = cudaStreamSynchronize(e->stream);
    if (st == cudaSuccess && refresh_ids(e)) st = cudaErrorUnknown;
    if (st == cudaSuccess) {
      for (int p = 0; p < e->np; p++) memcpy(lists + (size_t)e->h_ids[p] * cap, &h[(size_t)p * cap], sizeof(int) * cap);
      if (counts && get_ints(e, counts, e->P.nnodes, 1)) st = cudaErrorUnknown;
    }
  }
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

// ---- asynchronous results download (scheme call) ----------------------------------------------------------------
// snapshot: every field is transposed to the caller's AoS layout into its own region of `snap` on the compute stream;
// flush: the D2H copies run on dl_stream after the snapshot, while the compute stream is already stepping on.
static int snapshot_prepare(nlps_engine* e) {
  const int D = e->D, T = e->T, DD = D * D;
  const size_t per = 6 * (size_t)D + 6 * (size_t)T + DD + 11 + 2 + 4 + (e->alme.bten ? 8 : 0);  // aLME: metric + ellipsoid
  const size_t need = per * (size_t)std::max(e->np, 1) + 2 * 40;
  if (!e->dl_stream) {
    CUDA_OK(cudaStreamCreateWithFlags(&e->dl_stream, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreateWithFlags(&e->dl_ready, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&e->dl_done, cudaEventDisableTiming));
  }
  if (e->snap_doubles < need) {
    size_t fr = 0, tot = 0;
    cudaMemGetInfo(&fr, &tot);
    const size_t want = need + need / 8;  // slab populations grow by migration
    if (!use_pool() && want * sizeof(double) > fr / 2) return 1;  // not worth half of what is left: synchronous path
    if (dev_alloc(e, &e->snap, want)) { cudaGetLastError(); e->snap = nullptr; e->snap_doubles = 0; return 1; }
    e->snap_doubles = want;
  }
  e->snap_off = 0;
  e->dl_copies.clear();
  return 0;
}
static int snapshot_flush(nlps_engine* e) {
  CUDA_OK(cudaStreamWaitEvent(e->dl_stream, e->dl_ready, 0));
  for (const auto& c : e->dl_copies) CUDA_OK(cudaMemcpyAsync(c.h, c.d, c.bytes, cudaMemcpyDeviceToHost, e->dl_stream));
  CUDA_OK(cudaEventRecord(e->dl_done, e->dl_stream));
  CUDA_OK(cudaEventSynchronize(e->dl_done));
  e->dl_copies.clear();
  return 0;
}

int nlps_b200_run_async(nlps_engine* e, int first_step, int count) {
  cudaSetDevice(e->device);
  for (int k = first_step; k < first_step + count; k++)
    for (int s_ = NLPS_STAGE_SEARCH; s_ <= NLPS_STAGE_G2P; s_++) enqueue_stage(e, s_, k);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
int nlps_b200_sync(nlps_engine* e) {
  cudaSetDevice(e->device);
  return poll_error(e);
}
int nlps_b200_download_begin(nlps_engine* e, nlps_particles* out) {
  cudaSetDevice(e->device);
  if (e->dl_pending) return 1;  // one snapshot at a time
  if (e->slab_on || snapshot_prepare(e)) return nlps_b200_download(e, out);  // rows by global id need the host scatter
  e->dl_async = 1;
  const int rc = download_impl(e, out, 0);
  e->dl_async = 0;
  if (rc) return rc;
  CUDA_OK(cudaEventRecord(e->dl_ready, e->stream));
  e->dl_pending = 1;
  return 0;
}
int nlps_b200_download_end(nlps_engine* e) {
  cudaSetDevice(e->device);
  if (!e->dl_pending) return 0;
  e->dl_pending = 0;
  return snapshot_flush(e);
}

// Page-lock the caller's field buffers for the duration of a scheme call: the D2H copies before every results
// step then run at PCIe/C2C speed instead of through the driver's pageable bounce buffers.  Best effort.
static void pin_state(const nlps_particles* st, int D, int T, bool on, std::vector<void*>& pinned) {
  if (!on) {
    for (void* p : pinned) cudaHostUnregister(p);
    pinned.clear();
    return;
  }
  const size_t n = (size_t)st->n;
  auto reg = [&](const void* p, size_t bytes) {
    if (!p || bytes < (1u << 20)) return;
    if (cudaHostRegister((void*)p, bytes, cudaHostRegisterDefault) == cudaSuccess) pinned.push_back((void*)p);
    else cudaGetLastError();
  };
  const double* vec[] = {st->x_GC, st->dis, st->D_dis, st->vel, st->acc, st->lambda};
  for (const double* p : vec) reg(p, n * D * 8);
  const double* ten[] = {st->F_n, st->F_n1, st->DF, st->b_e_n, st->b_e_n1, st->Stress};
  for (const double* p : ten) reg(p, n * T * 8);
  reg(st->C_ep, n * D * D * 8);
  const double* sca[] = {st->J_n, st->J_n1, st->mass, st->rho, st->Vol_0, st->W, st->EPS_n, st->EPS_n1, st->Kappa_n, st->Kappa_n1, st->Beta};
  for (const double* p : sca) reg(p, n * 8);
}

static int scheme_call(const nlps_mesh* mesh, const nlps_solver* solver, int n_bounds, const nlps_load* bounds,
                       int n_neumann, const nlps_load* neumann, const double* gravity, int n_materials,
                       const nlps_material* materials, nlps_particles* state, const nlps_slab* slab, int* ids_out,
                       int run_initialize, int results_every, nlps_results_cb cb, void* user, int device) {
  char msg[256];
  const bool timing = getenv("NLPS_TIMING") != nullptr;
  auto now = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
  double t_create = now(), t_run = 0.0, t_io = 0.0;
  std::vector<void*> pinned;
  // registering costs about as much as ten pageable downloads (measured: +0.12 s / -0.022 s per download at 10^6
  // particles), so only long output sequences pay for it
  if (results_every > 0 && (solver->num_steps - solver->initial_step) / results_every >= 12 && !getenv("NLPS_NO_PIN"))
    pin_state(state, mesh->ndim, mesh->ndim == 2 ? 5 : 9, true, pinned);
  nlps_engine* e = create_any(mesh, solver, n_bounds, bounds, n_neumann, neumann, gravity, n_materials, materials, state,
                              slab, device, msg, sizeof(msg));
  if (!e) { pin_state(state, 0, 0, false, pinned); return 1; }
  t_create = now() - t_create;
  const bool compact = slab && slab->global_id;
  const int n_in = state->n;
  auto fetch = [&]() {
    if (!compact) return nlps_b200_download(e, state);
    if (e->np > n_in) {
      fprintf(stderr, "nlps_b200_u_verlet_slab: the slab now holds %d particles, the caller's buffers %d rows\n", e->np, n_in);
      return 1;
    }
    state->n = n_in;
    int rc = nlps_b200_download_local(e, state, ids_out);
    state->n = e->np;
    return rc;
  };
  int status = 0;
  if (run_initialize) status = nlps_b200_initialize_lme(e);
  int k = solver->initial_step;
  // Results steps: the fields of step k are snapshotted on the device, the next chunk of steps is enqueued, and only
  // then are the snapshot's D2H copies issued (on their own stream) and cb(k) called: copies and the caller's output
  // code overlap the stepping.  The caller's buffers hold step k when cb(k) runs and are not touched again before it
  // returns.  NLPS_SYNC_IO=1 restores download-then-continue.
  const bool async_io = !getenv("NLPS_SYNC_IO");
  bool have_snap = false;
  int snap_k = -1;
  auto snapshot = [&]() {  // returns 0 ok (have_snap set), 1 error; falls back to a synchronous fetch
    if (!async_io || snapshot_prepare(e)) return fetch();
    e->dl_async = 1;
    const int rc = fetch();
    e->dl_async = 0;
    if (rc) return rc;
    if (cudaEventRecord(e->dl_ready, e->stream) != cudaSuccess) return 1;
    have_snap = true;
    return 0;
  };
  while (!status && k < solver->num_steps) {
    int chunk = solver->num_steps - k;
    if (results_every > 0) {  // results after every step with TimeStep % ResultsTimeStep == 0 (U-Verlet.c:1097)
      int nxt = ((k + results_every - 1) / results_every) * results_every;
      chunk = std::min(chunk, nxt - k + 1);
    }
    double t0 = now();
    for (int q = k; q < k + chunk; q++)
      for (int s_ = NLPS_STAGE_SEARCH; s_ <= NLPS_STAGE_G2P; s_++) enqueue_stage(e, s_, q);
    if (have_snap) {
      double t1 = now();
      status = snapshot_flush(e);
      have_snap = false;
      if (!status && cb) cb(snap_k, user);
      t_io += now() - t1;
      t0 += now() - t1;
    }
    if (!status) status = poll_error(e);
    t_run += now() - t0;
    k += chunk;
    if (!status && results_every > 0 && ((k - 1) % results_every == 0)) {
      t0 = now();
      snap_k = k - 1;
      status = snapshot();
      if (!status && !have_snap && cb) cb(snap_k, user);  // synchronous fallback: the data is already there
      t_io += now() - t0;
    }
  }
  double t0 = now();
  if (!status && have_snap) {
    status = snapshot_flush(e);
    have_snap = false;
    if (!status && cb) cb(snap_k, user);
  }
  if (!status) {
    status = snapshot();
    if (!status && have_snap) status = snapshot_flush(e);
  }
  t_io += now() - t0;
  t0 = now();
  nlps_b200_destroy(e);
  pin_state(state, 0, 0, false, pinned);
  if (timing)
    fprintf(stderr, "nlps_b200 scheme call: create %.3f s, steps %.3f s, downloads %.3f s, destroy %.3f s\n", t_create, t_run, t_io,
            now() - t0);
  return status;
}

int nlps_b200_u_verlet(const nlps_mesh* mesh, const nlps_solver* solver, int n_bounds, const nlps_load* bounds,
                       int n_neumann, const nlps_load* neumann, const double* gravity, int n_materials,
                       const nlps_material* materials, nlps_particles* state, int run_initialize, int results_every,
                       nlps_results_cb cb, void* user, int device) {
  return scheme_call(mesh, solver, n_bounds, bounds, n_neumann, neumann, gravity, n_materials, materials, state, nullptr,
                     nullptr, run_initialize, results_every, cb, user, device);
}

int nlps_b200_u_verlet_slab(const nlps_mesh* mesh, const nlps_solver* solver, int n_bounds, const nlps_load* bounds,
                            int n_neumann, const nlps_load* neumann, const double* gravity, int n_materials,
                            const nlps_material* materials, nlps_particles* state, const nlps_slab* slab, int* ids_out,
                            int run_initialize, int results_every, nlps_results_cb cb, void* user, int device) {
  if (!slab) return 1;
  return scheme_call(mesh, solver, n_bounds, bounds, n_neumann, neumann, gravity, n_materials, materials, state, slab,
                     ids_out, run_initialize, results_every, cb, user, device);
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
  MatTable pt;  // the call's own table: no engine of the process is touched
  memset(&pt, 0, sizeof(pt));
  MatParams* hm = pt.m;
  hm[0] = MatParams{sm.type, sm.rho, sm.E, sm.nu, sm.reference_pressure, sm.kappa_0, sm.hardening_modulus,
                    sm.plastic_strain_0, sm.phi_frictional, sm.psi_frictional, sm.exponent_hardening_ortiz,
                    sm.cohesion, sm.alpha_hardening_borja, sm.a_hardening_borja[0], sm.a_hardening_borja[1],
                    sm.a_hardening_borja[2]};
  mat_hoist(hm[0], ndim);
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
  if (ndim == 2) k_stress_points<2><<<nblk(n, 64), 64>>>(n, pt, rp, dDF, dF1, dJ, dbe, deps, dkap, dS, dbe1, deps1, dkap1, dW, dC, dst);
  else k_stress_points<3><<<nblk(n, 64), 64>>>(n, pt, rp, dDF, dF1, dJ, dbe, deps, dkap, dS, dbe1, deps1, dkap1, dW, dC, dst);
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

#include "nlps_implicit.inl"
