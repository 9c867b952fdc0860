// nlps_cellwarp.cu -- warp-per-cell kernels of the explicit NPC-FS step in three dimensions (sm_100a, fp64).
// (2D decks run the block-per-cell-group kernels of nlps_engine.cu: 25 ring slots and 4 particles per cell leave
// nothing to split over the lanes of a warp.)
//
// One WARP owns one occupied cell at a time (a cell = all particles with the same closest node I0 = one contiguous run
// of the cell-sorted particle order) and walks a contiguous range of cells; nothing is shared between the warps of a
// block: no block barrier anywhere.  The warp stages the 2-ring node data of its cell in its slice of shared memory and
// takes the particles in chunks of 8; the neighbour loop of a particle is split over 4 lanes (8 .. 32 when the weights
// of a long list must wait in the compact cache), partial sums meet in xor-shuffles.
//
// Pipeline: the ring node ids of the NEXT cell and the cell record after it arrive by cp.async (LDGSTS) while the
// current cell computes, so that a cell starts with one dependent global round trip (node data) instead of three
// (record -> ring ids -> node data); the particle rows of the next cell are prefetched to L2.
//
// Neighbour lists: the LME kernel tests the 2-ring slots with the lanes interleaved (no bank conflicts), stores the
// reference's list as a bitmask and ALSO as ascending slot ids, one byte per neighbour (P.clist): every kernel of the
// step then runs over neighbours only (~40 of 125 ring slots at gamma = 6) with balanced lanes.
//
// Particle-to-grid without atomics.  Mass / momentum: the weights of the converged LME evaluation sit in a dense
// [particle][slot] table in shared memory (zero for non-neighbours); lane k sums column k over the particles of the
// chunk and writes record k of the cell's run part[(cell rank, slot)] straight from registers.  Forces: the particles
// of a chunk are taken one after the other, the lanes run over THAT particle's neighbours (distinct slots: no conflict)
// and add into the cell's accumulators in shared memory, which are flushed as the same kind of run.  The runs are
// CELL-major: 32-byte records, 4 KB contiguous per cell -- with the slot-major layout of the 2D kernels (coalesced for
// the reader) the 125 scattered stores of a cell were a third of the kinematics kernel (GridDev::part, DESIGN.md
// section 4).  The LME kernel also leaves the mask of the cell's non-zero records (G.cum = union of its particles'
// lists).  The node kernels k_grid_disp / k_grid_acc (nlps_engine.cu) gather the non-zero records of a node with a warp
// and add them in a fixed order: deterministic.
//
// The LME kernel leaves 1/Z and the inverse Hessian J^-1 of the converged evaluation per particle (P.zi, P.ji): the
// kinematics, force and G2P kernels evaluate the weights exp(-beta |l|^2 + lambda.l) once and need no second pass
// for Z, r, J.  Plastic laws run in a kernel of their own with a thread per particle (the return mapping would leave
// 3 of the 4 lanes of a particle idle): gather (DF) -> stress -> force sums.
//
// Reference citations are relative to nl-partsol/src of migmolper/NL-PartSol.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "nlps_cellwarp.h"

namespace {

constexpr int D = 3;            // every kernel of this file
constexpr int LPP0 = 4;         // lanes per particle of the base mapping
constexpr int PPW = 32 / LPP0;  // particles per chunk
constexpr int NJ = D * (D + 1) / 2;
constexpr int PVN = 2 * D + 2 + D * D + D;  // per-particle record of the force sums
constexpr unsigned FULL = 0xffffffffu;

struct Carve {
  size_t off = 0;
  __host__ __device__ size_t take(size_t bytes) { size_t o = off; off = (off + bytes + 15) & ~(size_t)15; return o; }
};
// shared-memory slice of ONE warp, evaluated identically on host (size) and device (offsets)
struct CwLayout {
  size_t mrec, um, ids, rank, q, X, U, A, acc, wd, list, wts, pv, total;
  __host__ __device__ CwLayout(const CwCfg& c, int kernel) {
    Carve k;
    const size_t SL = c.SL;
    const bool lme = kernel == CW_LME_P2G, fused = kernel == CW_KIN_FUSED, force = kernel == CW_FORCE;
    const bool wantQ = false;  // ranks and transposed-ring positions were the addresses of the slot-major cell sums
    const bool wantU = fused || kernel == CW_KIN_GATHER || kernel == CW_G2P;
    mrec = k.take(16 * 4);                    // ring of 4 cell records (int4)
    um = k.take(4 * MAX_MASK_WORDS);          // union of the neighbour masks of the cell's particles
    ids = k.take(4 * (size_t)c.CL);           // ring node ids of the next cell (cp.async)
    rank = wantQ ? k.take(4 * SL) : 0;
    q = wantQ ? k.take(SL) : 0;
    X = k.take(8 * D * SL);
    U = wantU ? k.take(8 * D * SL) : 0;
    A = kernel == CW_G2P ? k.take(8 * D * SL) : 0;
    acc = (fused || force) ? k.take(8 * D * SL) : 0;
    wd = lme ? k.take(8 * (size_t)PPW * (c.CL + 1)) : 0;   // dense weights [particle][slot], row stride CL + 1
    list = k.take((size_t)PPW * c.CL);
    wts = fused ? k.take(8 * (size_t)PPW * c.NC) : 0;
    pv = (fused || force) ? k.take(8 * (size_t)PPW * PVN) : (lme ? k.take(8 * (size_t)PPW * 4) : 0);
    total = k.off;
  }
};
struct WarpTile {
  int4* mrec;
  uint32_t* um;
  int *ids, *rank;
  unsigned char *q, *list;
  double *X, *U, *A, *acc, *wd, *wts, *pv;
};
__device__ __forceinline__ WarpTile carve_tile(unsigned char* w, const CwLayout& L) {
  WarpTile T;
  T.mrec = (int4*)(w + L.mrec); T.um = (uint32_t*)(w + L.um); T.ids = (int*)(w + L.ids); T.rank = (int*)(w + L.rank);
  T.q = w + L.q; T.list = w + L.list;
  T.X = (double*)(w + L.X); T.U = (double*)(w + L.U); T.A = (double*)(w + L.A);
  T.acc = (double*)(w + L.acc); T.wd = (double*)(w + L.wd); T.wts = (double*)(w + L.wts); T.pv = (double*)(w + L.pv);
  return T;
}

// ---- cp.async (LDGSTS): global -> shared without a register in between
__device__ __forceinline__ void cp_async4(void* sdst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// hint: the particle rows [t, t + 8) of `ncomp` SoA components (next cell of this warp) -> L2
__device__ __forceinline__ void cw_prefetch_rows(const double* f, int ld, int ncomp, int t, int np, int lane) {
  if (lane < 2 * ncomp) {
    const int tt = min(t + ((lane & 1) ? 7 : 0), np - 1);
    if (tt >= 0) prefetch_l2(f + (size_t)(lane >> 1) * ld + tt);
  }
}

// ---- cell pipeline.  Warp gw takes the cells gw, gw + nw, gw + 2 nw, ... (nw = warps of the grid): at any moment the
// warps of an SM -- and of the whole chip -- work on NEIGHBOURING cells, whose ring nodes they share in L1 / L2 (a
// contiguous range per warp scatters the active cells over the whole mesh: measured 20 % slower).  Cell records live in
// a ring of 4 (index i & 3 for the warp's i-th cell): at cell i the records of i and i + 1 are there, i + 2 is on its way;
// the ring node ids of cell i are in T.ids, those of i + 1 are requested as soon as the staging of i has read them.
struct Cell { int B, t0, t1, base, len; };
__device__ __forceinline__ void cw_pipe_init(const MeshDev& m, const GridDev& G, const WarpTile& T, int gw, int nw, int nocc,
                                             int lane) {
  if (gw >= nocc) return;
  if (lane < 2 && gw + lane * nw < nocc) T.mrec[lane] = G.occ_meta[gw + lane * nw];
  __syncwarp();
  const int4 r = T.mrec[0];
  for (int k = lane; k < (r.w & 511); k += 32) T.ids[k] = m.r2i[r.z + k];
  __syncwarp();
}
__device__ __forceinline__ Cell cw_cell(const WarpTile& T, int i) {
  const int4 r = T.mrec[i & 3];
  Cell c;
  c.B = r.x; c.t0 = r.y; c.base = r.z; c.len = r.w & 511;
  c.t1 = r.y + (r.w >> 9);
  return c;
}
// first particle slot of the warp's next cell (for the L2 prefetch of its rows), -1 when there is none
__device__ __forceinline__ int cw_next_t0(const WarpTile& T, int i, int u, int nw, int nocc) {
  return (u + nw < nocc) ? T.mrec[(i + 1) & 3].y : -1;
}
// call after the staging of cell i (= global cell u) has consumed T.ids (and a __syncwarp)
__device__ __forceinline__ void cw_pipe_next(const MeshDev& m, const GridDev& G, const WarpTile& T, int i, int u, int nw, int nocc,
                                             int lane) {
  if (u + nw < nocc) {
    const int4 r1 = T.mrec[(i + 1) & 3];
    for (int k = lane; k < (r1.w & 511); k += 32) cp_async4(&T.ids[k], &m.r2i[r1.z + k]);
  }
  if (lane == 0 && u + 2 * nw < nocc) cp_async16(&T.mrec[(i + 2) & 3], &G.occ_meta[u + 2 * nw]);
}
// end of a cell: everything requested for the next one has landed
__device__ __forceinline__ void cw_pipe_wait() {
  cp_async_wait_all();
  __syncwarp();
}

// 2-ring node data of the cell -> the warp's shared-memory slice (ring ids from T.ids).  A node record (32 bytes of
// coordinates, 32 + 32 bytes of nodal dU | a) is fetched by a PAIR of lanes, 16 bytes each: a warp instruction touches 16
// sectors and uses all of each (a lane per record would touch 32 sectors per instruction and half of each, and the L1
// data pipe is the busiest unit of these kernels); four records per lane pair in flight.
// UNION: only the slots that are a neighbour of some particle of the cell are fetched (T.um), ~70 of 125 at gamma = 6.
// SENT: inactive nodes and the padding of short rings get coordinates at 1e300, so that the neighbour test of the LME
// kernel rejects them without looking at the rank (x - 1e300 squared overflows to +inf, never <= s*)
template <bool WANT_Q, int NF, bool SENT, bool UNION>
__device__ __forceinline__ void cw_stage(const MeshDev& m, const GridDev& G, int SL, const Cell& c, const WarpTile& T, int lane) {
  constexpr int U = 4;
  const int half = lane & 1, row0 = lane >> 1;
  for (int k0 = 0; k0 < SL; k0 += 16 * U) {
    int node[U];
    bool need[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int k = k0 + 16 * u + row0;
      node[u] = (k < c.len) ? T.ids[k] : -1;
      need[u] = node[u] >= 0 && (!UNION || ((T.um[k >> 5] >> (k & 31)) & 1u));
    }
    int rank[U];
    unsigned char qv[U];
    double2 x[U], uu[U], aa[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int k = k0 + 16 * u + row0;
      rank[u] = -1;
      qv[u] = 0;
      if ((WANT_Q || SENT) && node[u] >= 0 && half == 0) rank[u] = G.arank[node[u]];
      if (WANT_Q && node[u] >= 0 && half == 0) qv[u] = m.r2q[c.base + k];
      if (need[u]) {
        x[u] = *reinterpret_cast<const double2*>(&m.X[(size_t)node[u] * NS<D>::X + 2 * half]);
        if (NF >= 1) uu[u] = *reinterpret_cast<const double2*>(&G.UA[(size_t)node[u] * 2 * NS<D>::X + 2 * half]);
        if (NF >= 2) aa[u] = *reinterpret_cast<const double2*>(&G.UA[(size_t)node[u] * 2 * NS<D>::X + NS<D>::X + 2 * half]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int k = k0 + 16 * u + row0;
      if (k < SL) {
        if (WANT_Q && half == 0) { T.rank[k] = rank[u]; T.q[k] = qv[u]; }
        double* dx = T.X + (size_t)k * D;
        if (SENT) {
          const int rk = __shfl_sync(__activemask(), rank[u], lane & ~1);
          const bool far = node[u] < 0 || rk < 0;
          if (half == 0) { dx[0] = far ? 1.0e300 : x[u].x; dx[1] = far ? 1.0e300 : x[u].y; }
          else dx[2] = far ? 1.0e300 : x[u].x;
        } else if (need[u]) {
          if (half == 0) { dx[0] = x[u].x; dx[1] = x[u].y; } else dx[2] = x[u].x;
        }
        if (NF >= 1 && need[u]) {
          double* du = T.U + (size_t)k * D;
          if (half == 0) { du[0] = uu[u].x; du[1] = uu[u].y; } else du[2] = uu[u].x;
        }
        if (NF >= 2 && need[u]) {
          double* da = T.A + (size_t)k * D;
          if (half == 0) { da[0] = aa[u].x; da[1] = aa[u].y; } else da[2] = aa[u].x;
        }
      }
    }
  }
}
// Union of the neighbour masks of the cell's particles -> T.um, as the LME kernel of this step left it: the slots of the
// cell that carry a non-zero mass sum (G.cum, one coalesced load instead of a particle-id and a mask round trip).
__device__ __forceinline__ void cw_union_cum(const GridDev& G, int u, int W, const WarpTile& T, int lane) {
  if (lane < MAX_MASK_WORDS) T.um[lane] = lane < W ? G.cum[(size_t)u * G.cm_w + lane] : 0u;
  __syncwarp();
}
// shape of a pass (fused kinematics kernel): the 8 particles of a chunk share 8 * NC compact-cache entries; lists
// longer than NC take the entries of 2, 4 or 8 particle slots and the chunk is then worked off in 2, 4 or 8 passes with
// 8, 16 or 32 lanes per particle
__device__ __forceinline__ void cw_pass_shape(int n_mine, int NC, int& stride, int& nper, int& lshift) {
  int nmax = n_mine;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(FULL, nmax, o));
  stride = max(4, (nmax + 3) & ~3);
  nper = PPW;
  lshift = 2;  // log2(lanes per particle)
  while (nper > 1 && nper * stride > PPW * NC) { nper >>= 1; lshift++; }
}
// neighbour bitmask -> ascending slot ids.  Lane `sub` of the particle's 4 lanes writes the bits b with b % 4 == sub of
// every mask word (neighbours cluster in a few words of the chain-ordered ring: equal RANGES of bits would leave most
// lanes idle), at the position the bit has in ascending slot order.
template <int W>
__device__ __forceinline__ void cw_compact(const uint32_t (&mk)[W], int sub, unsigned char* lst) {
  const uint32_t pat = 0x11111111u << sub;
  int pre = 0;
#pragma unroll
  for (int w = 0; w < W; w++) {
    const uint32_t mw = mk[w];
    uint32_t mine = mw & pat;
    while (mine) {
      const int b = __ffs(mine) - 1;
      lst[pre + __popc(mw & ((1u << b) - 1u))] = (unsigned char)(32 * w + b);
      mine &= mine - 1;
    }
    pre += __popc(mw);
  }
}
// spread the 8 low bits of x to the bit positions 0, 4, 8, ... 28
__device__ __forceinline__ uint32_t spread8(uint32_t x) {
  x &= 0xffu;
  x = (x | (x << 12)) & 0x000F000Fu;
  x = (x | (x << 6)) & 0x03030303u;
  x = (x | (x << 3)) & 0x11111111u;
  return x;
}
// the particle's compact list P.clist -> the warp's slice (4-byte words, the particle's lanes share the copy)
__device__ __forceinline__ void cw_fetch_list(const unsigned char* gl, unsigned char* sl, int n, int subp, int lpp) {
  const uint32_t* g = reinterpret_cast<const uint32_t*>(gl);
  uint32_t* s_ = reinterpret_cast<uint32_t*>(sl);
  for (int i = subp; i < (n + 3) >> 2; i += lpp) s_[i] = g[i];
}
template <int Dd>
__device__ __forceinline__ void sym_to_full(const double* s, double* A) {
  if (Dd == 2) { A[0] = s[0]; A[1] = s[1]; A[2] = s[1]; A[3] = s[2]; }
  else { A[0] = s[0]; A[1] = s[1]; A[2] = s[2]; A[3] = s[1]; A[4] = s[3]; A[5] = s[4]; A[6] = s[2]; A[7] = s[4]; A[8] = s[5]; }
}
template <int Dd>
__device__ __forceinline__ void full_to_sym(const double* A, double* s) {
  if (Dd == 2) { s[0] = A[0]; s[1] = A[1]; s[2] = A[3]; }
  else { s[0] = A[0]; s[1] = A[1]; s[2] = A[2]; s[3] = A[4]; s[4] = A[5]; s[5] = A[8]; }
}

// ---------------------------------------------------------------------------
// K0 + K1: tributary__LME__ (LME.c:1019-1099) with the PREVIOUS beta, beta__LME__ (LME.c:177-185),
// __lambda_Newton_Rapson (LME.c:272-353, warm start), __predictor_PARTICLES (U-Verlet.c:229-253) and the cell sums
// of the lumped mass and of the mass-weighted displacement increment (U-Verlet.c:166-225, 301-367).
template <int W>
__global__ void __launch_bounds__(128, 4) cw_lme_p2g(const MeshDev m, const PartDev P, const GridDev G, const StepParams sp,
                                                     const CwCfg cfg, int* err, int do_predictor) {
  extern __shared__ __align__(16) unsigned char smem[];
  // exp() with the 2^(j/32) table: the Newton loop of this kernel sits at the register limit of 4 blocks per SM and
  // the table version needs 8 registers less than the polynomial one of the other kernels
  __shared__ double s_tab[32];
  if (threadIdx.x < 32) s_tab[threadIdx.x] = g_exp2tab[threadIdx.x];
  __syncthreads();  // the only block-wide barrier of the kernel
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const CwLayout L(cfg, CW_LME_P2G);
  const WarpTile T = carve_tile(smem + (size_t)wib * L.total, L);
  constexpr int NV = 1 + D;
  const int SL = cfg.SL, ld = P.ld, CL = cfg.CL, WS = cfg.CL + 1;  // WS: row stride of the dense weight table
  const int nocc = *G.n_occ;
  const int nwarps = gridDim.x * wpb, gw = blockIdx.x * wpb + wib;
  for (int e = lane; e < PPW * WS; e += 32) T.wd[e] = 0.0;  // invariant: the table is all zero between chunks
  cw_pipe_init(m, G, T, gw, nwarps, nocc, lane);
  const int j = lane >> 2, sub = lane & (LPP0 - 1);  // particle of the chunk, lane of the particle
  unsigned char* const lst = T.list + j * CL;
  double* const wd = T.wd + (size_t)j * WS;
  double* const pw = T.pv + j * 4;  // per particle: m / Z and m / Z * DU_p
  for (int u = gw, ic = 0; u < nocc; u += nwarps, ic++) {
    const Cell c = cw_cell(T, ic);
    const int tn = cw_next_t0(T, ic, u, nwarps, nocc);
    if (tn >= 0) {  // the first rows of the next cell -> L2 while this cell computes
      cw_prefetch_rows(P.x, ld, D, tn, P.np, lane);
      cw_prefetch_rows(P.lam, ld, D, tn, P.np, lane);
      cw_prefetch_rows(do_predictor ? P.vel : P.ddis, ld, D, tn, P.np, lane);
      if (do_predictor) cw_prefetch_rows(P.acc, ld, D, tn, P.np, lane);
      cw_prefetch_rows(P.beta, ld, 1, tn, P.np, lane);
      cw_prefetch_rows(P.mass, ld, 1, tn, P.np, lane);
      cw_prefetch_rows(P.sstar, ld, 1, tn, P.np, lane);
    }
    cw_stage<false, 0, true, false>(m, G, SL, c, T, lane);  // (no ranks / ring positions: the cell sums are stored cell-major)
    __syncwarp();
    cw_pipe_next(m, G, T, ic, u, nwarps, nocc, lane);
    for (int tb = c.t0; tb < c.t1; tb += PPW) {
      const int t = tb + j;
      const bool valid = t < c.t1;
      int p = 0;
      double xp[D], lam[D], beta_old = 1.0, sstar = -1.0;
#pragma unroll
      for (int i = 0; i < D; i++) { xp[i] = 0.0; lam[i] = 0.0; }
      if (valid) {
        p = G.plist[t];
        beta_old = P.beta[p];
        sstar = P.sstar[p];
        const double mp = P.mass[p];
        double dd[D];
#pragma unroll
        for (int i = 0; i < D; i++) {
          xp[i] = P.x[i * ld + p];
          lam[i] = P.lam[i * ld + p];
          const double v = do_predictor ? P.vel[i * ld + p] : (sp.proj ? sp.proj[i * ld + p] : P.ddis[i * ld + p]);
          if (do_predictor) {  // gamma = 0.5, U-Verlet.c:76,248
            const double a = P.acc[i * ld + p];
            dd[i] = sp.dt * v + 0.5 * (sp.dt * sp.dt) * a;
            if (sub == 0) {
              P.ddis[i * ld + p] = dd[i];
              P.vel[i * ld + p] = v + (1 - 0.5) * sp.dt * a;
            }
          } else {
            dd[i] = v;
          }
        }
        if (sub == 0) { pw[0] = mp; pw[1] = mp * dd[0]; pw[2] = mp * dd[1]; pw[3] = mp * dd[2]; }  // times 1 / Z below
      }
      uint32_t mk[W];
      if (sp.reuse_lists) {
#pragma unroll
        for (int w = 0; w < W; w++) mk[w] = valid ? P.mask[(size_t)w * ld + p] : 0u;
      } else {
        // lane `sub` tests the slots 4 i + sub of the cell's 2-ring (neighbouring lanes read neighbouring rows of the
        // tile: no bank conflicts): distances rounded exactly as the reference does (no FMA); s <= s* <=> sqrt(s) <= Ra
        // (LME.c:1052,1074) with s* = s*(beta of the previous step), kept per particle (P.sstar)
        constexpr int WL = (W + 3) / 4;  // 32-bit words of tested bits per lane (bit i <-> slot 4 i + sub)
        uint32_t part[WL];
#pragma unroll
        for (int w = 0; w < WL; w++) part[w] = 0u;
        if (valid) {
          for (int k = sub; k < c.len; k += LPP0) {
            double l[D];
            const double s = dist2_exact<D>(xp, T.X + k * D, l);
            if (s <= sstar) {
              if constexpr (WL == 1) part[0] |= 1u << (k >> 2);
              else part[k >> 7] |= 1u << ((k >> 2) & 31);
            }
          }
        }
        // mask word w = slots [32 w, 32 w + 32) = bits [8 w, 8 w + 8) of the four lanes, interleaved
#pragma unroll
        for (int w = 0; w < W; w++) {
          uint32_t v = spread8(part[w / 4] >> (8 * (w % 4))) << sub;
          v |= __shfl_xor_sync(FULL, v, 1);
          v |= __shfl_xor_sync(FULL, v, 2);
          mk[w] = v;
        }
        if (valid) {
#pragma unroll
          for (int w = 0; w < W; w++)
            if ((w & (LPP0 - 1)) == sub) P.mask[(size_t)w * ld + p] = mk[w];
        }
      }
      int n = 0;
#pragma unroll
      for (int w = 0; w < W; w++) n += __popc(mk[w]);
      bool ok = valid;
      if (valid && !sp.reuse_lists && sub == 0) P.nnodes[p] = n;
      if (valid && n < D + 1) { if (sub == 0) latch_error(err, NLPS_ERR_FEW_NEIGHBOURS, P.orig[p]); ok = false; }
      double beta = beta_old;
      if (valid && !sp.reuse_lists) {
        const double h = m.h_avg[c.B];
        beta = __ddiv_rn(sp.gamma_lme, __dmul_rn(h, h));
        if (sub == 0) { P.beta[p] = beta; P.sstar[p] = m.sst[c.B]; }
      }
      // the list as ascending slot ids: shared memory for this kernel, P.clist for the other kernels of the step
      if (valid) cw_compact<W>(mk, sub, lst);
      __syncwarp();
      if (valid) {
        const uint32_t* sl = reinterpret_cast<const uint32_t*>(lst);
        uint32_t* gl = reinterpret_cast<uint32_t*>(P.clist + (size_t)p * CL);
        for (int i = sub; i < (n + 3) >> 2; i += LPP0) gl[i] = sl[i];
      }
      // ---- Newton on lambda: lane `sub` takes the neighbour pairs (2 sub, 2 sub + 1), + 8, ...
      const int n_ = ok ? n : 0;
      int NumIter = 0;
      bool act = ok;
      double Zi = 0.0;
      while (__any_sync(FULL, act)) {
        double Z = 0.0, r[D], JJ[D * D];
#pragma unroll
        for (int i = 0; i < D; i++) r[i] = 0.0;
#pragma unroll
        for (int i = 0; i < D * D; i++) JJ[i] = 0.0;
        if (act) {
          for (int i0 = 2 * sub; i0 < n_; i0 += 2 * LPP0) {
            const bool two = i0 + 1 < n_;
            const unsigned kk = *reinterpret_cast<const unsigned short*>(lst + i0);
            const int k0 = kk & 0xffu, k1 = two ? (int)(kk >> 8) : k0;
            double l0[D], l1[D], ll0 = 0.0, lx0 = 0.0, ll1 = 0.0, lx1 = 0.0;
#pragma unroll
            for (int i = 0; i < D; i++) {
              l0[i] = xp[i] - T.X[k0 * D + i];
              l1[i] = xp[i] - T.X[k1 * D + i];
              ll0 += l0[i] * l0[i];
              ll1 += l1[i] * l1[i];
              lx0 += l0[i] * lam[i];
              lx1 += l1[i] * lam[i];
            }
            const double e0 = fexp(-beta * ll0 + lx0, s_tab);
            const double e1 = two ? fexp(-beta * ll1 + lx1, s_tab) : 0.0;
            wd[k0] = e0;  // the weights of the LAST evaluation are the ones the cell sums use
            if (two) wd[k1] = e1;
            Z += e0;
#pragma unroll
            for (int i = 0; i < D; i++) {
              const double el = e0 * l0[i];
              r[i] += el;
#pragma unroll
              for (int jj = i; jj < D; jj++) JJ[i * D + jj] += el * l0[jj];
            }
            Z += e1;
#pragma unroll
            for (int i = 0; i < D; i++) {
              const double el = e1 * l1[i];
              r[i] += el;
#pragma unroll
              for (int jj = i; jj < D; jj++) JJ[i * D + jj] += el * l1[jj];
            }
          }
        }
#pragma unroll
        for (int o = 1; o < LPP0; o <<= 1) {
          Z += __shfl_xor_sync(FULL, Z, o);
#pragma unroll
          for (int i = 0; i < D; i++) {
            r[i] += __shfl_xor_sync(FULL, r[i], o);
#pragma unroll
            for (int jj = i; jj < D; jj++) JJ[i * D + jj] += __shfl_xor_sync(FULL, JJ[i * D + jj], o);
          }
        }
        if (act) {
          Zi = 1.0 / Z;
          double nr = 0.0;
#pragma unroll
          for (int i = 0; i < D; i++) { r[i] *= Zi; nr += r[i] * r[i]; }
          nr = sqrt(nr);
#pragma unroll
          for (int i = 0; i < D; i++)
#pragma unroll
            for (int jj = i; jj < D; jj++) {
              JJ[i * D + jj] = JJ[i * D + jj] * Zi - r[i] * r[jj];
              JJ[jj * D + i] = JJ[i * D + jj];
            }
          double Ji[D * D];
          if (nr > sp.tol_wrapper) {
            if (rcond_as_reference<D>(JJ) < 1E-8) {
              ok = false;
              act = false;
              if (sub == 0) latch_error(err, NLPS_ERR_SINGULAR_HESSIAN, P.orig[p]);
            } else {
              inverse<D>(JJ, Ji);
#pragma unroll
              for (int i = 0; i < D; i++) {
                double dl = 0.0;
#pragma unroll
                for (int jj = 0; jj < D; jj++) dl += Ji[i * D + jj] * r[jj];
                lam[i] -= dl;
              }
              NumIter++;
              act = NumIter <= sp.max_iter_lme;
            }
          } else {  // converged: this evaluation's Z and Hessian are the step's shape-function data
            inverse<D>(JJ, Ji);
            if (sub == 0) {
              double Jis[NJ];
              full_to_sym<D>(Ji, Jis);
#pragma unroll
              for (int i = 0; i < NJ; i++) P.ji[(size_t)i * ld + p] = Jis[i];
            }
            act = false;
          }
        }
      }
      if (ok && NumIter >= sp.max_iter_lme && sub == 0) latch_error(err, NLPS_ERR_NEWTON_LME, P.orig[p]);
      if (valid && sub == 0) {
#pragma unroll
        for (int i = 0; i < D; i++) P.lam[i * ld + p] = lam[i];
        const double z = ok ? Zi : 0.0;
        P.zi[p] = z;
#pragma unroll
        for (int i = 0; i < 4; i++) pw[i] *= z;
      }
      __syncwarp();
      // ---- cell sums: lane k sums column k of the weight table over the particles of the chunk
      const int cnt = min(PPW, c.t1 - tb);
      const bool first = tb == c.t0;
      // (cell-major partial sums: the records of a cell are one contiguous run, lane k writes record k)
      for (int k = lane; k < c.len; k += 32) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (int jq = 0; jq < cnt; jq++) {
          const double w = T.wd[(size_t)jq * WS + k];
          const double2 q01 = *reinterpret_cast<const double2*>(T.pv + jq * 4);
          const double2 q23 = *reinterpret_cast<const double2*>(T.pv + jq * 4 + 2);
          a0 += w * q01.x; a1 += w * q01.y; a2 += w * q23.x; a3 += w * q23.y;
        }
        {
          double* dst = G.part + ((size_t)u * G.cm_sl + k) * NV;
          if (!first) {  // cells with more than 8 particles: add to what the earlier chunks wrote
            const double2 o01 = *reinterpret_cast<const double2*>(dst), o23 = *reinterpret_cast<const double2*>(dst + 2);
            a0 += o01.x; a1 += o01.y; a2 += o23.x; a3 += o23.y;
          }
          *reinterpret_cast<double2*>(dst) = make_double2(a0, a1);
          *reinterpret_cast<double2*>(dst + 2) = make_double2(a2, a3);
        }
        // which records of the cell are non-zero (a slot is some particle's neighbour <=> its mass sum is positive): the
        // node kernels skip the others
        const uint32_t nz = __ballot_sync(__activemask(), a0 != 0.0);
        if ((k & 31) == 0) G.cum[(size_t)u * G.cm_w + (k >> 5)] = nz;
      }
      __syncwarp();
      // the table returns to zero: every lane clears the entries it wrote
      for (int i0 = 2 * sub; i0 < n_; i0 += 2 * LPP0) {
        const unsigned kk = *reinterpret_cast<const unsigned short*>(lst + i0);
        wd[kk & 0xffu] = 0.0;
        if (i0 + 1 < n_) wd[kk >> 8] = 0.0;
      }
      __syncwarp();
    }
    cw_pipe_wait();
  }
}

// ---------------------------------------------------------------------------
// Particle part of K2: F_n1 = DF F_n (compute-Strains.c:76-105), J > 0 (U-Verlet.c:608-613), rho /= det DF
// (U-Verlet.c:630-632), stress (Constitutive.c:18-258) and the force operator G = V0 tau DF^-T J^-1, so that
// f_A = sum_p N_A (G_p l_A + t_p) == -V0 tau (DF^-T gradN_A) + N_A T A0 (U-Newmark-beta.c:1257-1374 with
// Shape-Functions.c:405-448).  MAT: compile-time law of a uniform cloud, -1 = per particle.
template <int D, int MAT>
__device__ __forceinline__ bool particle_stress(const PartDev& P, const StepParams& sp, const MatTable& mt, int p,
                                                const double* DF, const double* Ji, int* err, double* Gp) {
  constexpr int T = (D == 2) ? 5 : 9;
  const int ld = P.ld;
  double Fn[D * D], Fn1[D * D];
#pragma unroll
  for (int i = 0; i < D * D; i++) Fn[i] = P.F_n[(size_t)i * ld + p];
  const double rho_p = P.rho[p], V0 = P.vol0[p];
  const int mid = P.matidx[p];
  const MatParams& mat = mt.m[mid];
  const int mtype = (MAT >= 0) ? MAT : mat.type;
  double be[T], eps = 0.0, kap = 0.0, back[3] = {0.0, 0.0, 0.0};
  if (MAT < 0 && P.back) {
#pragma unroll
    for (int i = 0; i < 3; i++) back[i] = P.back[(size_t)i * ld + p];
  }
  if (mtype != NLPS_MAT_NEO_HOOKEAN_WRIGGERS) {
#pragma unroll
    for (int i = 0; i < T; i++) be[i] = P.be_n[(size_t)i * ld + p];
    eps = P.eps_n[p];
    kap = P.kap_n[p];
  }
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int jj = 0; jj < D; jj++) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++) s += DF[i * D + k] * Fn[k * D + jj];
      Fn1[i * D + jj] = s;
      P.F_n1[(size_t)(i * D + jj) * ld + p] = s;
    }
  const double J1 = det<D>(Fn1);
  P.J_n1[p] = J1;
#pragma unroll
  for (int i = 0; i < D * D; i++) Gp[i] = 0.0;
  if (J1 <= 0.0) { latch_error(err, NLPS_ERR_NEGATIVE_JACOBIAN, P.orig[p]); return false; }
  const double dJ = det<D>(DF);
  if (!sp.implicit) P.rho[p] = rho_p / dJ;  // the implicit scheme updates rho once, after convergence
  double tau[T], Wp = 0.0;
  if (mtype == NLPS_MAT_NEO_HOOKEAN_WRIGGERS) {
    stress_neo_hookean<D>(mat, Fn1, J1, tau, Wp);
  } else {
    double cep[D * D];
    int st;
    if (MAT == NLPS_MAT_DRUCKER_PRAGER) st = stress_drucker_prager<D>(mat, sp.rp, DF, be, eps, kap, tau, Wp, cep);
    else if (MAT == NLPS_MAT_MATSUOKA_NAKAI) st = stress_matsuoka_nakai<D>(mat, sp.rp, DF, be, eps, kap, tau, Wp, cep);
    else st = stress_with_history<D>(mtype, mat, sp.rp, DF, Fn1, be, eps, kap, back, tau, Wp, cep);
    if (st != 0) { latch_error(err, st, P.orig[p]); return false; }
    if (MAT >= 0 || mat_has_history(mtype)) {
#pragma unroll
      for (int i = 0; i < T; i++) P.be_n1[(size_t)i * ld + p] = be[i];
      P.eps_n1[p] = eps;
      if (MAT >= 0 || mtype != NLPS_MAT_VON_MISES) P.kap_n1[p] = kap;  // Von-Mises never touches Kappa
      if (MAT < 0 && mtype == NLPS_MAT_VON_MISES && P.back) {
#pragma unroll
        for (int i = 0; i < 3; i++) P.back[(size_t)i * ld + p] = back[i];
      }
      if (sp.rp.want_cep)
#pragma unroll
        for (int i = 0; i < D * D; i++) P.cep[(size_t)i * ld + p] = cep[i];
    }
  }
#pragma unroll
  for (int i = 0; i < T; i++) P.stress[(size_t)i * ld + p] = tau[i];
  P.W[p] = Wp;
  double DFi[D * D];
  const double dd = inverse<D>(DF, DFi);
  if (dd == 0.0) { latch_error(err, NLPS_ERR_SINGULAR_DF, P.orig[p]); return false; }
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
  return true;
}

// thread per particle (plastic and mixed clouds): coalesced SoA, all lanes busy in the return mapping
template <int D, int MAT>
__global__ void __launch_bounds__(128) cw_stress(const PartDev P, const StepParams sp, const __grid_constant__ MatTable mt, int* err) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.np) return;
  constexpr int NJ = D * (D + 1) / 2;
  const int ld = P.ld;
  double DF[D * D], Js[NJ], Ji[D * D], Gp[D * D];
#pragma unroll
  for (int i = 0; i < D * D; i++) DF[i] = P.DF[(size_t)i * ld + p];
#pragma unroll
  for (int i = 0; i < NJ; i++) Js[i] = P.ji[(size_t)i * ld + p];
  sym_to_full<D>(Js, Ji);
  particle_stress<D, MAT>(P, sp, mt, p, DF, Ji, err, Gp);
#pragma unroll
  for (int i = 0; i < D * D; i++) P.gop[(size_t)i * ld + p] = Gp[i];
}

// ---------------------------------------------------------------------------
// K2 + K3.  MODE CW_KIN_FUSED (Neo-Hookean clouds): DF = I + sum_A DU_A (x) gradN_A with gradN_a = -p_a J^-1 l_a
// (compute-Strains.c:20-44, LME.c:836-891), the particle part above, and the cell sums of the nodal forces with the
// weights of the gather still in shared memory.  CW_KIN_GATHER: DF only.  CW_FORCE: force sums from P.gop.
// Neighbours come from the compact lists of the LME kernel (P.clist, P.nnodes), 1 / Z and J^-1 from P.zi, P.ji.
template <int MODE>
__global__ void __launch_bounds__(128, MODE == CW_KIN_FUSED ? 3 : 4) cw_kin(const MeshDev m, const PartDev P, const GridDev G, const StepParams sp,
                                                 const CwCfg cfg, const __grid_constant__ MatTable mt, int* err,
                                                 int has_traction) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const CwLayout L(cfg, MODE);
  const WarpTile T = carve_tile(smem + (size_t)wib * L.total, L);
  constexpr bool GATHER = MODE != CW_FORCE, SCATTER = MODE != CW_KIN_GATHER, FUSED = MODE == CW_KIN_FUSED;
  // per-particle record of the force sums: x (D) | lambda (D) | beta | 1/Z | G (D*D) | t (D)
  constexpr int PX = 0, PL = D, PB = 2 * D, PZ = 2 * D + 1, PG = 2 * D + 2, PT = 2 * D + 2 + D * D;
  const int SL = cfg.SL, ld = P.ld, CL = cfg.CL;
  const int nocc = *G.n_occ;
  const int nwarps = gridDim.x * wpb, gw = blockIdx.x * wpb + wib;
  cw_pipe_init(m, G, T, gw, nwarps, nocc, lane);
  for (int u = gw, ic = 0; u < nocc; u += nwarps, ic++) {
    const Cell c = cw_cell(T, ic);
    const int tn = cw_next_t0(T, ic, u, nwarps, nocc);
    if (tn >= 0) {  // the first rows of the next cell -> L2 while this cell computes
      cw_prefetch_rows(P.x, ld, D, tn, P.np, lane);
      cw_prefetch_rows(P.lam, ld, D, tn, P.np, lane);
      cw_prefetch_rows(P.beta, ld, 1, tn, P.np, lane);
      cw_prefetch_rows(P.zi, ld, 1, tn, P.np, lane);
      if (GATHER) cw_prefetch_rows(P.ji, ld, NJ, tn, P.np, lane);
      if (FUSED) {
        cw_prefetch_rows(P.F_n, ld, D * D, tn, P.np, lane);
        cw_prefetch_rows(P.rho, ld, 1, tn, P.np, lane);
        cw_prefetch_rows(P.vol0, ld, 1, tn, P.np, lane);
      }
      if (MODE == CW_FORCE) cw_prefetch_rows(P.gop, ld, D * D, tn, P.np, lane);
      if (lane < PPW && tn + lane < P.np) prefetch_l2(P.clist + (size_t)(tn + lane) * CL);
      if (lane >= 16 && lane < 16 + cfg.W) prefetch_l2(&P.mask[(size_t)(lane - 16) * ld + tn]);  // the union of the masks opens the next cell
    }
    cw_union_cum(G, u, cfg.W, T, lane);
    cw_stage<false, GATHER ? 1 : 0, false, true>(m, G, SL, c, T, lane);
    if (SCATTER)
      for (int e = lane; e < SL * D; e += 32) T.acc[e] = 0.0;
    __syncwarp();
    cw_pipe_next(m, G, T, ic, u, nwarps, nocc, lane);
    for (int tb = c.t0; tb < c.t1; tb += PPW) {
      const int t = tb + (lane >> 2);
      const bool valid = t < c.t1;
      int p = 0, n = 0;
      if (valid) {
        p = G.plist[t];
        n = P.nnodes[p];
      }
      int stride = 0, nper = PPW, lshift = 2;
      if (FUSED) cw_pass_shape(n, cfg.NC, stride, nper, lshift);  // the weights of the gather wait in the compact cache
      const int LPP = 1 << lshift;
      for (int h = 0; h < PPW / nper; h++) {
        const int jp = lane >> lshift, subp = lane & (LPP - 1);
        const int src = (h * nper + jp) * LPP0;
        const int p_ = __shfl_sync(FULL, p, src), n_ = __shfl_sync(FULL, n, src);
        const bool here = __shfl_sync(FULL, (int)valid, src) != 0;
        unsigned char* lst = T.list + jp * CL;
        double* wt = T.wts + jp * stride;
        double* pvj = T.pv + jp * PVN;
        double x_[D], lam_[D], beta_ = 0.0, Zi = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) { x_[i] = 0.0; lam_[i] = 0.0; }
        if (here) {
          cw_fetch_list(P.clist + (size_t)p_ * CL, lst, n_, subp, LPP);
#pragma unroll
          for (int i = 0; i < D; i++) { x_[i] = P.x[i * ld + p_]; lam_[i] = P.lam[i * ld + p_]; }
          beta_ = P.beta[p_];
          Zi = P.zi[p_];
        }
        __syncwarp();
        if (GATHER) {
          double Bm[D * D];
#pragma unroll
          for (int i = 0; i < D * D; i++) Bm[i] = 0.0;
          for (int i0 = 2 * subp; i0 < n_; i0 += 2 * LPP) {
            const bool two = i0 + 1 < n_;
            const unsigned kk = *reinterpret_cast<const unsigned short*>(lst + i0);
            const int k0 = kk & 0xffu, k1 = two ? (int)(kk >> 8) : k0;
            double l0[D], l1[D], ll0 = 0.0, lx0 = 0.0, ll1 = 0.0, lx1 = 0.0;
#pragma unroll
            for (int i = 0; i < D; i++) {
              l0[i] = x_[i] - T.X[k0 * D + i];
              l1[i] = x_[i] - T.X[k1 * D + i];
              ll0 += l0[i] * l0[i];
              ll1 += l1[i] * l1[i];
              lx0 += l0[i] * lam_[i];
              lx1 += l1[i] * lam_[i];
            }
            const double e0 = fexp_poly(-beta_ * ll0 + lx0);
            const double e1 = two ? fexp_poly(-beta_ * ll1 + lx1) : 0.0;
            if (FUSED) {
              wt[i0] = e0;
              if (two) wt[i0 + 1] = e1;
            }
#pragma unroll
            for (int i = 0; i < D; i++) {
              const double eu0 = e0 * T.U[k0 * D + i], eu1 = e1 * T.U[k1 * D + i];
#pragma unroll
              for (int jj = 0; jj < D; jj++) Bm[i * D + jj] += eu0 * l0[jj];
#pragma unroll
              for (int jj = 0; jj < D; jj++) Bm[i * D + jj] += eu1 * l1[jj];
            }
          }
          for (int o = 1; o < LPP; o <<= 1) {
#pragma unroll
            for (int i = 0; i < D * D; i++) Bm[i] += __shfl_xor_sync(FULL, Bm[i], o);
          }
          if (here && subp == 0) {
            double Js[NJ], Ji[D * D], DF[D * D];
#pragma unroll
            for (int i = 0; i < NJ; i++) Js[i] = P.ji[(size_t)i * ld + p_];
            sym_to_full<D>(Js, Ji);
#pragma unroll
            for (int i = 0; i < D; i++)
#pragma unroll
              for (int jj = 0; jj < D; jj++) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < D; k++) s += Bm[i * D + k] * Ji[k * D + jj];
                DF[i * D + jj] = ((i == jj) ? 1.0 : 0.0) - s * Zi;
                P.DF[(size_t)(i * D + jj) * ld + p_] = DF[i * D + jj];
              }
            if (FUSED) {
              double Gp[D * D];
              const bool ok = particle_stress<D, NLPS_MAT_NEO_HOOKEAN_WRIGGERS>(P, sp, mt, p_, DF, Ji, err, Gp);
#pragma unroll
              for (int i = 0; i < D; i++) {
                pvj[PX + i] = x_[i];
                pvj[PT + i] = (ok && has_traction) ? P.trac[(size_t)i * ld + p_] : 0.0;
              }
              pvj[PZ] = ok ? Zi : 0.0;
#pragma unroll
              for (int i = 0; i < D * D; i++) pvj[PG + i] = Gp[i];
            }
          }
        } else {  // CW_FORCE: the record comes from the stress kernel
          if (here && subp == 0) {
#pragma unroll
            for (int i = 0; i < D; i++) {
              pvj[PX + i] = x_[i];
              pvj[PL + i] = lam_[i];
              pvj[PT + i] = has_traction ? P.trac[(size_t)i * ld + p_] : 0.0;
            }
            pvj[PB] = beta_;
            pvj[PZ] = Zi;
#pragma unroll
            for (int i = 0; i < D * D; i++) pvj[PG + i] = P.gop[(size_t)i * ld + p_];
          }
        }
        if (SCATTER) {
          __syncwarp();
          // the particles of the pass one after the other, lanes over that particle's neighbours
          for (int jq = 0; jq < nper; jq++) {
            const int nq = __shfl_sync(FULL, here ? n_ : 0, jq << lshift);
            if (nq == 0) continue;  // uniform
            const double* pq = T.pv + jq * PVN;
            double xq[D], Gq[D * D], tq[D], lq_[D], bq = 0.0;
            const double zq = pq[PZ];
#pragma unroll
            for (int i = 0; i < D; i++) { xq[i] = pq[PX + i]; tq[i] = pq[PT + i]; lq_[i] = 0.0; }
#pragma unroll
            for (int i = 0; i < D * D; i++) Gq[i] = pq[PG + i];
            if (MODE == CW_FORCE) {
              bq = pq[PB];
#pragma unroll
              for (int i = 0; i < D; i++) lq_[i] = pq[PL + i];
            }
            const unsigned char* lq = T.list + jq * CL;
            const double* wtq = T.wts + jq * stride;
            for (int i = lane; i < nq; i += 32) {
              const int k = lq[i];
              double l[D], ll = 0.0, lx = 0.0;
#pragma unroll
              for (int d = 0; d < D; d++) {
                l[d] = xq[d] - T.X[k * D + d];
                if (MODE == CW_FORCE) { ll += l[d] * l[d]; lx += l[d] * lq_[d]; }
              }
              const double e = (MODE == CW_FORCE) ? fexp_poly(-bq * ll + lx) : wtq[i];
              const double N = e * zq;
#pragma unroll
              for (int d = 0; d < D; d++) {
                double gl = tq[d];
#pragma unroll
                for (int kk = 0; kk < D; kk++) gl += Gq[d * D + kk] * l[kk];
                T.acc[d * SL + k] += N * gl;
              }
            }
            __syncwarp();
          }
        }
        __syncwarp();
      }
    }
    if (SCATTER) {
      // cell-major partial sums: 32-byte records, the cell's run is contiguous (see GridDev::part)
      for (int k = lane; k < c.len; k += 32) {
        double* dst = G.part + ((size_t)u * G.cm_sl + k) * 4;
        *reinterpret_cast<double2*>(dst) = make_double2(T.acc[k], T.acc[SL + k]);
        *reinterpret_cast<double2*>(dst + 2) = make_double2(T.acc[2 * SL + k], 0.0);
      }
    }
    cw_pipe_wait();
  }
}

// ---------------------------------------------------------------------------
// K4: G2P + corrector (U-Verlet.c:963-1084): a_p = sum N_A a_A, DU_p = sum N_A DU_A, v += gamma dt a, x += DU,
// dis += DU.  The n+1 -> n roll of F, J, b_e, kappa, EPS is a pointer swap on the host side of the engine.
__global__ void __launch_bounds__(128, 4) cw_g2p(const MeshDev m, const PartDev P, const GridDev G, const StepParams sp,
                                                 const CwCfg cfg) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const CwLayout L(cfg, CW_G2P);
  const WarpTile T = carve_tile(smem + (size_t)wib * L.total, L);
  const int SL = cfg.SL, ld = P.ld, CL = cfg.CL;
  const int nocc = *G.n_occ;
  const int nwarps = gridDim.x * wpb, gw = blockIdx.x * wpb + wib;
  cw_pipe_init(m, G, T, gw, nwarps, nocc, lane);
  // 4 lanes per particle, lane `sub` takes the neighbour pairs (2 sub, 2 sub + 1), + 8, ...
  const int j = lane >> 2, sub = lane & (LPP0 - 1);
  unsigned char* const lst = T.list + j * CL;
  for (int u = gw, ic = 0; u < nocc; u += nwarps, ic++) {
    const Cell c = cw_cell(T, ic);
    const int tn = cw_next_t0(T, ic, u, nwarps, nocc);
    if (tn >= 0) {
      cw_prefetch_rows(P.x, ld, D, tn, P.np, lane);
      cw_prefetch_rows(P.lam, ld, D, tn, P.np, lane);
      cw_prefetch_rows(P.vel, ld, D, tn, P.np, lane);
      cw_prefetch_rows(P.dis, ld, D, tn, P.np, lane);
      cw_prefetch_rows(P.beta, ld, 1, tn, P.np, lane);
      cw_prefetch_rows(P.zi, ld, 1, tn, P.np, lane);
      if (lane < PPW && tn + lane < P.np) prefetch_l2(P.clist + (size_t)(tn + lane) * CL);
      if (lane >= 16 && lane < 16 + cfg.W) prefetch_l2(&P.mask[(size_t)(lane - 16) * ld + tn]);  // the union of the masks opens the next cell
    }
    cw_union_cum(G, u, cfg.W, T, lane);
    cw_stage<false, 2, false, true>(m, G, SL, c, T, lane);
    __syncwarp();
    cw_pipe_next(m, G, T, ic, u, nwarps, nocc, lane);
    for (int tb = c.t0; tb < c.t1; tb += PPW) {
      const int t = tb + j;
      const bool valid = t < c.t1;
      int p = 0, n = 0;
      double x_[D], lam_[D], a[D], du[D], beta_ = 0.0, Zi = 0.0, vel_s = 0.0, dis_s = 0.0;
#pragma unroll
      for (int i = 0; i < D; i++) { x_[i] = 0.0; lam_[i] = 0.0; a[i] = 0.0; du[i] = 0.0; }
      if (valid) {
        p = G.plist[t];
        n = P.nnodes[p];
        cw_fetch_list(P.clist + (size_t)p * CL, lst, n, sub, LPP0);
#pragma unroll
        for (int i = 0; i < D; i++) { x_[i] = P.x[i * ld + p]; lam_[i] = P.lam[i * ld + p]; }
        beta_ = P.beta[p];
        Zi = P.zi[p];
        // component `sub` of the particle is updated by lane `sub`: its old values are requested now, used after the loop
        if (sub < D) { vel_s = P.vel[sub * ld + p]; dis_s = P.dis[sub * ld + p]; }
      }
      __syncwarp();
      for (int i0 = 2 * sub; i0 < n; i0 += 2 * LPP0) {
        const bool two = i0 + 1 < n;
        const unsigned kk = *reinterpret_cast<const unsigned short*>(lst + i0);
        const int k0 = kk & 0xffu, k1 = two ? (int)(kk >> 8) : k0;
        double ll0 = 0.0, lx0 = 0.0, ll1 = 0.0, lx1 = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) {
          const double l0 = x_[i] - T.X[k0 * D + i], l1 = x_[i] - T.X[k1 * D + i];
          ll0 += l0 * l0;
          ll1 += l1 * l1;
          lx0 += l0 * lam_[i];
          lx1 += l1 * lam_[i];
        }
        const double e0 = fexp_poly(-beta_ * ll0 + lx0);
        const double e1 = two ? fexp_poly(-beta_ * ll1 + lx1) : 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) {
          a[i] += e0 * T.A[k0 * D + i];
          du[i] += e0 * T.U[k0 * D + i];
          a[i] += e1 * T.A[k1 * D + i];
          du[i] += e1 * T.U[k1 * D + i];
        }
      }
#pragma unroll
      for (int o = 1; o < LPP0; o <<= 1) {
#pragma unroll
        for (int i = 0; i < D; i++) {
          a[i] += __shfl_xor_sync(FULL, a[i], o);
          du[i] += __shfl_xor_sync(FULL, du[i], o);
        }
      }
      if (valid && sub < D) {
        double as = a[0], ds = du[0], xs = x_[0];
#pragma unroll
        for (int i = 1; i < D; i++)
          if (sub == i) { as = a[i]; ds = du[i]; xs = x_[i]; }
        const double ai = as * Zi, di = ds * Zi;
        P.acc[sub * ld + p] = ai;
        P.ddis[sub * ld + p] = di;
        P.vel[sub * ld + p] = vel_s + 0.5 * sp.dt * ai;
        P.x[sub * ld + p] = xs + di;
        P.dis[sub * ld + p] = dis_s + di;
      }
      __syncwarp();
    }
    cw_pipe_wait();
  }
}

// ---------------------------------------------------------------------------
template <class K>
int cw_prepare(K kernel, int id, const CwLaunch& L, CwState& st, int sm_count, int max_smem_optin) {
  if (st.ready[id]) return 0;
  const size_t per_warp = CwLayout(L.cfg, id).total;
  st.smem[id] = per_warp * L.cfg.warps;
  // (the kernels also hold 256 bytes of static shared memory: ask for what the slices need, not for the device limit)
  if (st.smem[id] + 1024 > (size_t)max_smem_optin ||
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st.smem[id]) != cudaSuccess) {
    fprintf(stderr, "nlps_b200: cannot reserve %zu bytes of shared memory for the warp tiles: %s\n", st.smem[id],
            cudaGetErrorString(cudaGetLastError()));
    return 1;
  }
  int nb = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, L.cfg.warps * 32, st.smem[id]) != cudaSuccess) return 1;
  st.grid[id] = sm_count * std::max(1, nb);
  if (const char* s_ = getenv("NLPS_CW_GRID")) st.grid[id] = std::max(1, atoi(s_));
  st.ready[id] = 1;
  return 0;
}

}  // namespace

size_t cw_smem_bytes(int kernel, const CwCfg& cfg) { return CwLayout(cfg, kernel).total * cfg.warps; }

int cw_launch_lme_p2g(int ndim, int W, const CwLaunch& L, CwState& st, int sm_count, int max_smem_optin, int do_predictor) {
  if (ndim != 3) return 1;
#define GO_(w)                                                                                                        \
  {                                                                                                                   \
    auto kf = cw_lme_p2g<w>;                                                                                          \
    if (cw_prepare(kf, CW_LME_P2G, L, st, sm_count, max_smem_optin)) return 1;                                         \
    kf<<<std::min(L.max_blocks, st.grid[CW_LME_P2G]), L.cfg.warps * 32, st.smem[CW_LME_P2G], L.stream>>>(              \
        L.m, L.P, L.G, L.sp, L.cfg, L.err, do_predictor);                                                             \
    return 0;                                                                                                         \
  }
  if (W == 4) GO_(4)
  if (W == 8) GO_(8)
#undef GO_
  return 1;
}

int cw_launch_kin(int ndim, int mode, const CwLaunch& L, CwState& st, int sm_count, int max_smem_optin, const MatTable& mt,
                  int has_traction) {
  if (ndim != 3) return 1;
#define GO_(md)                                                                                                       \
  {                                                                                                                   \
    auto kf = cw_kin<md>;                                                                                             \
    if (cw_prepare(kf, md, L, st, sm_count, max_smem_optin)) return 1;                                                 \
    kf<<<std::min(L.max_blocks, st.grid[md]), L.cfg.warps * 32, st.smem[md], L.stream>>>(L.m, L.P, L.G, L.sp, L.cfg,   \
                                                                                         mt, L.err, has_traction);    \
    return 0;                                                                                                         \
  }
  if (mode == CW_KIN_FUSED) GO_(CW_KIN_FUSED)
  if (mode == CW_KIN_GATHER) GO_(CW_KIN_GATHER)
  if (mode == CW_FORCE) GO_(CW_FORCE)
#undef GO_
  return 1;
}

int cw_launch_stress(int ndim, const CwLaunch& L, const MatTable& mt, int uniform_mat) {
  if (ndim != 3) return 1;
  const int np = L.P.np;
  if (np <= 0) return 0;
  const int grid = (np + 127) / 128;
#define GO_(mat) { cw_stress<D, mat><<<grid, 128, 0, L.stream>>>(L.P, L.sp, mt, L.err); return 0; }
  switch (uniform_mat) { case 0: GO_(0) case 1: GO_(1) case 2: GO_(2) default: GO_(-1) }
#undef GO_
  return 1;
}

int cw_launch_g2p(int ndim, const CwLaunch& L, CwState& st, int sm_count, int max_smem_optin) {
  if (ndim != 3) return 1;
  auto kf = cw_g2p;
  if (cw_prepare(kf, CW_G2P, L, st, sm_count, max_smem_optin)) return 1;
  kf<<<std::min(L.max_blocks, st.grid[CW_G2P]), L.cfg.warps * 32, st.smem[CW_G2P], L.stream>>>(L.m, L.P, L.G, L.sp, L.cfg);
  return 0;
}
