// nlps_cellwarp.cu -- warp-per-cell kernels of the explicit NPC-FS step (sm_100a, fp64).
//
// One WARP owns a run of CPW consecutive occupied cells (a cell = all particles with the same closest node I0, i.e.
// one contiguous run of the cell-sorted particle order) and nothing is shared between the warps of a block: no block
// barrier anywhere, every warp is at its own point of its own cell.  The warp stages the 2-ring node data of its
// cells in its slice of shared memory and walks the particles in chunks of 8; the neighbour loop of a particle is
// split over 4 lanes (8 .. 32 for lists longer than the compact cache), partial sums meet in xor-shuffles.  A
// particle's neighbour list is COMPACTED by the LME kernel from its bitmask into ascending slot ids (P.clist, one byte
// per neighbour), so that every lane of every kernel of the step runs over neighbours only (3D: ~40 of 125 ring slots)
// and the lanes of a particle stay balanced.
//
// Particle-to-grid without atomics: the particles of a chunk are taken one after the other, the lanes of the warp
// run over THAT particle's neighbours (distinct slots: no conflict) and add into the warp's (cell, slot) accumulators
// in shared memory; a cell's accumulators go to part[(slot of the cell in the node's transposed ring, node rank)] and
// are summed per node in a fixed order by k_grid_disp / k_grid_acc (nlps_engine.cu).  Deterministic.
//
// The LME kernel leaves 1/Z and the inverse Hessian J^-1 of the converged evaluation per particle (P.zi, P.ji): the
// kinematics, force and G2P kernels of the same step evaluate the weights exp(-beta |l|^2 + lambda.l) once and need
// no second pass for Z, r, J.  Plastic laws run in a kernel of their own with a thread per particle (the return
// mapping would leave 3 of the 4 lanes of a particle idle): gather (DF) -> stress -> force sums.
//
// Reference citations are relative to nl-partsol/src of migmolper/NL-PartSol.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "nlps_cellwarp.h"

namespace {

constexpr int LPP0 = 4;         // lanes per particle of the base mapping
constexpr int PPW = 32 / LPP0;  // particles per chunk
constexpr unsigned FULL = 0xffffffffu;

template <int D> struct PV { static constexpr int N = 2 * D + 2 + D * D + D; };  // per-particle record of the force sums

struct Carve {
  size_t off = 0;
  __host__ __device__ size_t take(size_t bytes) { size_t o = off; off = (off + bytes + 15) & ~(size_t)15; return o; }
};
// shared-memory slice of ONE warp, evaluated identically on host (size) and device (offsets)
struct CwLayout {
  size_t meta, rank, q, X, U, A, acc, list, wts, pv, total;
  __host__ __device__ CwLayout(const CwCfg& c, int D, int kernel) {
    Carve k;
    const size_t pairs = (size_t)c.CPW * c.SL;
    const bool wantQ = kernel == CW_LME_P2G || kernel == CW_KIN_FUSED || kernel == CW_FORCE;
    const bool wantU = kernel == CW_KIN_FUSED || kernel == CW_KIN_GATHER || kernel == CW_G2P;
    const int nacc = kernel == CW_LME_P2G ? 1 + D : ((kernel == CW_KIN_FUSED || kernel == CW_FORCE) ? D : 0);
    const bool wantW = kernel == CW_LME_P2G || kernel == CW_KIN_FUSED;
    const bool wantPV = kernel == CW_KIN_FUSED || kernel == CW_FORCE;
    // LME kernel: x (D) | lambda (D) | DU_p (D) | beta | mass per particle, then p, cell, n (or -1) and the mask words
    const size_t pvk = kernel == CW_LME_P2G ? 8 * (size_t)PPW * (3 * D + 2) + 4 * (size_t)PPW * (3 + MAX_MASK_WORDS) : 0;
    meta = k.take(4 * (size_t)(4 * c.CPW + 1));
    rank = k.take(4 * pairs);
    q = wantQ ? k.take(pairs) : 0;
    X = k.take(8 * D * pairs);
    U = wantU ? k.take(8 * D * pairs) : 0;
    A = kernel == CW_G2P ? k.take(8 * D * pairs) : 0;
    acc = nacc ? k.take(8 * (size_t)nacc * pairs) : 0;
    list = kernel == CW_LME_P2G ? k.take((size_t)PPW * c.NC) : 0;
    wts = wantW ? k.take(8 * (size_t)PPW * c.NC) : 0;
    pv = wantPV ? k.take(8 * (size_t)PPW * (2 * D + 2 + D * D + D)) : (pvk ? k.take(pvk) : 0);
    total = k.off;
  }
};
struct WarpTile {
  int *cs, *base, *len, *B, *rank;
  unsigned char *q, *list;
  double *X, *U, *A, *acc, *wts, *pv;
};
__device__ __forceinline__ WarpTile carve_tile(unsigned char* w, const CwLayout& L, const CwCfg& c) {
  WarpTile T;
  T.cs = (int*)(w + L.meta); T.base = T.cs + (c.CPW + 1); T.len = T.base + c.CPW; T.B = T.len + c.CPW;
  T.rank = (int*)(w + L.rank); T.q = w + L.q; T.list = w + L.list;
  T.X = (double*)(w + L.X); T.U = (double*)(w + L.U); T.A = (double*)(w + L.A);
  T.acc = (double*)(w + L.acc); T.wts = (double*)(w + L.wts); T.pv = (double*)(w + L.pv);
  return T;
}

// metadata of the warp's cells: one record per lane (first particle slot, node, 2-ring base and length)
__device__ __forceinline__ void cw_meta(const GridDev& G, int nocc, int np, int c0, int ncell, const WarpTile& T, int lane) {
  if (lane <= ncell) {
    if (c0 + lane < nocc) {
      const int4 mt = G.occ_meta[c0 + lane];
      T.cs[lane] = mt.y;
      if (lane < ncell) { T.B[lane] = mt.x; T.base[lane] = mt.z; T.len[lane] = mt.w; }
    } else {
      T.cs[lane] = np;  // the last occupied cell ends at the last particle
    }
  }
  __syncwarp();
}
// 2-ring node data of the warp's cells -> its shared-memory slice; four (cell, slot) pairs per lane in flight
// SENT: inactive nodes and the padding of short rings get coordinates at 1e300, so that the neighbour test of the LME
// kernel rejects them without looking at the rank (x - 1e300 squared overflows to +inf, never <= s*)
template <int D, bool WANT_Q, int NF, bool SENT = false>
__device__ __forceinline__ void cw_stage(const MeshDev& m, const GridDev& G, const CwCfg& cfg, int ncell, const WarpTile& T,
                                         int lane) {
  constexpr int U = 4;
  const int SL = cfg.SL, npairs = ncell * SL;
  for (int e0 = lane; e0 < npairs; e0 += U * 32) {
    int node[U], idx[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int e = e0 + u * 32;
      node[u] = -1;
      idx[u] = 0;
      if (e < npairs) {
        const int c = (int)(((unsigned)e * cfg.magic) >> 21), k = e - c * SL;
        if (k < T.len[c]) { idx[u] = T.base[c] + k; node[u] = m.r2i[idx[u]]; }
      }
    }
    int rank[U];
    unsigned char qv[U];
    double2 x0[U], x1[U], u0[U], u1[U], a0[U], a1[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int nd = max(node[u], 0);  // invalid pairs load node 0 (harmless): unconditional independent loads
      const double* px = &m.X[(size_t)nd * NS<D>::X];
      const double* pu = &G.UA[(size_t)nd * 2 * NS<D>::X];
      rank[u] = G.arank[nd];
      x0[u] = *reinterpret_cast<const double2*>(px);
      if (D == 3) x1[u] = *reinterpret_cast<const double2*>(px + 2);
      if (WANT_Q) qv[u] = m.r2q[idx[u]];
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
      const int e = e0 + u * 32;
      if (e < npairs) {
        T.rank[e] = (node[u] >= 0) ? rank[u] : -1;
        if (WANT_Q) T.q[e] = qv[u];
        double* dx = T.X + (size_t)e * D;
        const bool far = SENT && (node[u] < 0 || rank[u] < 0);
        dx[0] = far ? 1.0e300 : x0[u].x; dx[1] = far ? 1.0e300 : x0[u].y;
        if (D == 3) dx[2] = far ? 1.0e300 : x1[u].x;
        if (NF >= 1) {
          double* du = T.U + (size_t)e * D;
          du[0] = u0[u].x; du[1] = u0[u].y;
          if (D == 3) du[2] = u1[u].x;
        }
        if (NF >= 2) {
          double* da = T.A + (size_t)e * D;
          da[0] = a0[u].x; da[1] = a0[u].y;
          if (D == 3) da[2] = a1[u].x;
        }
      }
    }
  }
}
__device__ __forceinline__ int cw_cell_of(const int* cs, int ncell, int t) {
  int c = 0;
  for (int i = 1; i < ncell; i++) c += (cs[i] <= t) ? 1 : 0;  // cs ascending; empty cells do not occur in the occupied list
  return c;
}
// shape of a pass: the 8 particles of a chunk share 8 * NC compact-cache entries; lists longer than NC take the
// entries of 2, 4 or 8 particle slots and the chunk is then worked off in 2, 4 or 8 passes with 8, 16 or 32 lanes per particle
__device__ __forceinline__ void cw_pass_shape(int n_mine, int NC, int& stride, int& nper, int& lshift) {
  int nmax = n_mine;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(FULL, nmax, o));
  stride = max(4, (nmax + 3) & ~3);
  nper = PPW;
  lshift = 2;  // log2(lanes per particle)
  while (nper > 1 && nper * stride > PPW * NC) { nper >>= 1; lshift++; }
}
// neighbour bitmask -> ascending slot ids.  Lane `subp` of the particle's LPP lanes writes the bits b with b % LPP ==
// subp of every mask word (neighbours cluster in a few words of the chain-ordered ring: equal RANGES of bits would leave
// most lanes idle), at the position the bit has in ascending slot order.
template <int W>
__device__ __forceinline__ void cw_compact(const uint32_t (&mk)[W], int lshift, int subp, unsigned char* lst) {
  // bits b = subp (mod LPP): LPP = 4 -> 0x11111111 << subp, 8 -> 0x01010101 << subp, 16 -> 0x00010001 << subp, 32 -> 1 << subp
  const uint32_t pat = (lshift == 2 ? 0x11111111u : (lshift == 3 ? 0x01010101u : (lshift == 4 ? 0x00010001u : 1u))) << subp;
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
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// hint: the particle rows [t, t + 8) of `ncomp` SoA components (next unit of this warp) -> L2
__device__ __forceinline__ void cw_prefetch_rows(const double* f, int ld, int ncomp, int t, int np, int lane) {
  if (lane < 2 * ncomp) {
    const int tt = min(t + ((lane & 1) ? 7 : 0), np - 1);
    if (tt >= 0) prefetch_l2(f + (size_t)(lane >> 1) * ld + tt);
  }
}
template <int D>
__device__ __forceinline__ void sym_to_full(const double* s, double* A) {
  if (D == 2) { A[0] = s[0]; A[1] = s[1]; A[2] = s[1]; A[3] = s[2]; }
  else { A[0] = s[0]; A[1] = s[1]; A[2] = s[2]; A[3] = s[1]; A[4] = s[3]; A[5] = s[4]; A[6] = s[2]; A[7] = s[4]; A[8] = s[5]; }
}
template <int D>
__device__ __forceinline__ void full_to_sym(const double* A, double* s) {
  if (D == 2) { s[0] = A[0]; s[1] = A[1]; s[2] = A[3]; }
  else { s[0] = A[0]; s[1] = A[1]; s[2] = A[2]; s[3] = A[4]; s[4] = A[5]; s[5] = A[8]; }
}

// ---------------------------------------------------------------------------
// K0 + K1: tributary__LME__ (LME.c:1019-1099) with the PREVIOUS beta, beta__LME__ (LME.c:177-185),
// __lambda_Newton_Rapson (LME.c:272-353, warm start), __predictor_PARTICLES (U-Verlet.c:229-253) and the cell sums
// of the lumped mass and of the mass-weighted displacement increment (U-Verlet.c:166-225, 301-367).
template <int D, int W>
__global__ void __launch_bounds__(128, 4) cw_lme_p2g(const MeshDev m, const PartDev P, const GridDev G, const StepParams sp,
                                                     const CwCfg cfg, int* err, int do_predictor) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double s_tab[32];
  if (threadIdx.x < 32) s_tab[threadIdx.x] = g_exp2tab[threadIdx.x];
  __syncthreads();  // the only block-wide barrier of the kernel
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const CwLayout L(cfg, D, CW_LME_P2G);
  const WarpTile T = carve_tile(smem + (size_t)wib * L.total, L, cfg);
  constexpr int NV = 1 + D, NJ = D * (D + 1) / 2;
  const int SL = cfg.SL, ld = P.ld;
  const int nocc = *G.n_occ, nunits = (nocc + cfg.CPW - 1) / cfg.CPW;
  // every warp walks a CONTIGUOUS range of units: its next cell is the neighbour of this one (shared ring nodes are
  // still in L1 / L2) and the particle rows it will need next follow the current ones
  const int nwarps = gridDim.x * wpb, gw = blockIdx.x * wpb + wib;
  const int u_lo = (int)((long long)gw * nunits / nwarps), u_hi = (int)((long long)(gw + 1) * nunits / nwarps);
  for (int u = u_lo; u < u_hi; u++) {
    const int c0 = u * cfg.CPW, ncell = min(cfg.CPW, nocc - c0);
    cw_meta(G, nocc, P.np, c0, ncell, T, lane);
    const int t0 = T.cs[0], t1 = T.cs[ncell];
    if (u + 1 < u_hi) {  // the first rows of the next unit -> L2 while this unit computes
      cw_prefetch_rows(P.x, ld, D, t1, P.np, lane);
      cw_prefetch_rows(P.lam, ld, D, t1, P.np, lane);
      cw_prefetch_rows(do_predictor ? P.vel : P.ddis, ld, D, t1, P.np, lane);
      if (do_predictor) cw_prefetch_rows(P.acc, ld, D, t1, P.np, lane);
      cw_prefetch_rows(P.beta, ld, 1, t1, P.np, lane);
      cw_prefetch_rows(P.mass, ld, 1, t1, P.np, lane);
      cw_prefetch_rows(P.sstar, ld, 1, t1, P.np, lane);
    }
    cw_stage<D, true, 0, true>(m, G, cfg, ncell, T, lane);
    for (int e = lane; e < ncell * SL * NV; e += 32) T.acc[e] = 0.0;
    __syncwarp();
    for (int tb = t0; tb < t1; tb += PPW) {
      // ---- base mapping: 4 lanes per particle; neighbour list and beta
      const int sub = lane & (LPP0 - 1);
      const int t = tb + (lane >> 2);
      const bool valid = t < t1;
      int p = 0, ci = 0, len = 0;
      double xp[D], lam[D], dd[D], beta_old = 1.0, mp = 0.0, sstar = -1.0;
#pragma unroll
      for (int i = 0; i < D; i++) { xp[i] = 0.0; lam[i] = 0.0; dd[i] = 0.0; }
      if (valid) {
        p = G.plist[t];
        ci = cw_cell_of(T.cs, ncell, t);
        len = T.len[ci];
        beta_old = P.beta[p];
        sstar = P.sstar[p];
        mp = P.mass[p];
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
        const double* Xc = T.X + (size_t)ci * SL * D;
        for (int k = sub; k < len; k += LPP0) {
          double l[D];
          const double s = dist2_exact<D>(xp, Xc + k * D, l);
          if (s <= sstar) {
            if constexpr (WL == 1) part[0] |= 1u << (k >> 2);
            else part[k >> 7] |= 1u << ((k >> 2) & 31);
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
        const int B = T.B[ci];
        const double h = m.h_avg[B];
        beta = __ddiv_rn(sp.gamma_lme, __dmul_rn(h, h));
        if (sub == 0) { P.beta[p] = beta; P.sstar[p] = m.sst[B]; }
      }
      int stride, nper, lshift;
      cw_pass_shape(ok ? n : 0, cfg.NC, stride, nper, lshift);
      const int LPP = 1 << lshift;
      // the particle's variables travel through shared memory: the base-mapping registers die here
      constexpr int PVK = 3 * D + 2;
      int* const pvi = (int*)(T.pv + PPW * PVK);
      if (sub == 0) {
        double* pj = T.pv + (lane >> 2) * PVK;
#pragma unroll
        for (int i = 0; i < D; i++) { pj[i] = xp[i]; pj[D + i] = lam[i]; pj[2 * D + i] = dd[i]; }
        pj[3 * D] = beta;
        pj[3 * D + 1] = mp;
        int* ij = pvi + (lane >> 2) * (3 + W);
        ij[0] = p; ij[1] = ci; ij[2] = valid ? (ok ? n : -1 - n) : -(1 << 20);  // n | too few neighbours | no particle
#pragma unroll
        for (int w = 0; w < W; w++) ij[3 + w] = (int)mk[w];
      }
      __syncwarp();
      for (int h = 0; h < PPW / nper; h++) {
        // ---- pass mapping: LPP lanes per particle
        const int jp = lane >> lshift, subp = lane & (LPP - 1);
        const int js = h * nper + jp;
        const double* pj = T.pv + js * PVK;
        const int* ij = pvi + js * (3 + W);
        double x_[D], lam_[D];
#pragma unroll
        for (int i = 0; i < D; i++) { x_[i] = pj[i]; lam_[i] = pj[D + i]; }
        const double beta_ = pj[3 * D];
        const int p_ = ij[0], ci_ = ij[1], n_raw = ij[2];
        const bool here = n_raw > -(1 << 20);
        bool ok_ = n_raw >= 0;
        const int n_ = ok_ ? n_raw : 0;
        const int n_list = here ? (ok_ ? n_raw : -1 - n_raw) : 0;  // the list is written for every particle (its consumers read P.nnodes entries)
        unsigned char* lst = T.list + jp * stride;
        double* wt = T.wts + jp * stride;
        {
          uint32_t mk_[W];
#pragma unroll
          for (int w = 0; w < W; w++) mk_[w] = (uint32_t)ij[3 + w];
          if (here) cw_compact<W>(mk_, lshift, subp, lst);
        }
        __syncwarp();
        // the compact list of the step -> P.clist (read by the kinematics, force and G2P kernels)
        if (here) {
          const uint32_t* sl = reinterpret_cast<const uint32_t*>(lst);
          uint32_t* gl = reinterpret_cast<uint32_t*>(P.clist + (size_t)p_ * cfg.CL);
          for (int i = subp; i < (n_list + 3) >> 2; i += LPP) gl[i] = sl[i];
        }
        // ---- Newton on lambda: lane subp takes the neighbour pairs (2 subp, 2 subp + 1), + 2 LPP, ...
        const double* Xc_ = T.X + (size_t)ci_ * SL * D;
        int NumIter = 0;
        bool act = ok_;
        double Zi = 0.0;
        while (__any_sync(FULL, act)) {
          double Z = 0.0, r[D], JJ[D * D];
#pragma unroll
          for (int i = 0; i < D; i++) r[i] = 0.0;
#pragma unroll
          for (int i = 0; i < D * D; i++) JJ[i] = 0.0;
          if (act) {
            for (int i0 = 2 * subp; i0 < n_; i0 += 2 * LPP) {
              const bool two = i0 + 1 < n_;
              const unsigned kk = *reinterpret_cast<const unsigned short*>(lst + i0);
              const int k0 = kk & 0xffu, k1 = two ? (int)(kk >> 8) : k0;
              double l0[D], l1[D], ll0 = 0.0, lx0 = 0.0, ll1 = 0.0, lx1 = 0.0;
#pragma unroll
              for (int i = 0; i < D; i++) {
                l0[i] = x_[i] - Xc_[k0 * D + i];
                l1[i] = x_[i] - Xc_[k1 * D + i];
                ll0 += l0[i] * l0[i];
                ll1 += l1[i] * l1[i];
                lx0 += l0[i] * lam_[i];
                lx1 += l1[i] * lam_[i];
              }
              const double e0 = fexp(-beta_ * ll0 + lx0, s_tab);
              const double e1 = two ? fexp(-beta_ * ll1 + lx1, s_tab) : 0.0;
              wt[i0] = e0;
              if (two) wt[i0 + 1] = e1;
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
          for (int o = 1; o < LPP; o <<= 1) {
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
                ok_ = false;
                act = false;
                if (subp == 0) latch_error(err, NLPS_ERR_SINGULAR_HESSIAN, P.orig[p_]);
              } else {
                inverse<D>(JJ, Ji);
#pragma unroll
                for (int i = 0; i < D; i++) {
                  double dl = 0.0;
#pragma unroll
                  for (int jj = 0; jj < D; jj++) dl += Ji[i * D + jj] * r[jj];
                  lam_[i] -= dl;
                }
                NumIter++;
                act = NumIter <= sp.max_iter_lme;
              }
            } else {  // converged: this evaluation's Z and Hessian are the step's shape-function data
              inverse<D>(JJ, Ji);
              if (subp == 0) {
                double Jis[NJ];
                full_to_sym<D>(Ji, Jis);
#pragma unroll
                for (int i = 0; i < NJ; i++) P.ji[(size_t)i * ld + p_] = Jis[i];
              }
              act = false;
            }
          }
        }
        if (ok_ && NumIter >= sp.max_iter_lme && subp == 0) latch_error(err, NLPS_ERR_NEWTON_LME, P.orig[p_]);
        if (here && subp == 0) {
#pragma unroll
          for (int i = 0; i < D; i++) P.lam[i * ld + p_] = lam_[i];
          P.zi[p_] = ok_ ? Zi : 0.0;
        }
        // ---- cell sums: the particles of the pass one after the other, lanes over that particle's neighbours
        const double wgt = ok_ ? pj[3 * D + 1] * Zi : 0.0;
        for (int jq = 0; jq < nper; jq++) {
          const int sl = jq << lshift;
          const int nq = __shfl_sync(FULL, ok_ ? n_ : 0, sl);
          const double wq = __shfl_sync(FULL, wgt, sl);
          double dq[D];
#pragma unroll
          for (int i = 0; i < D; i++) dq[i] = T.pv[(h * nper + jq) * PVK + 2 * D + i] * wq;
          const int cq = __shfl_sync(FULL, ci_, sl);
          const unsigned char* lq = T.list + jq * stride;
          const double* wtq = T.wts + jq * stride;
          double* accq = T.acc + (size_t)cq * NV * SL;
          for (int i = lane; i < nq; i += 32) {
            const int k = lq[i];
            const double e = wtq[i];
            accq[k] += e * wq;
#pragma unroll
            for (int d = 0; d < D; d++) accq[(1 + d) * SL + k] += e * dq[d];
          }
          __syncwarp();
        }
      }
    }
    // ---- the cells' sums -> part[(slot of the cell in the node's transposed ring, node rank)]
    for (int e = lane; e < ncell * SL; e += 32) {
      const int rank = T.rank[e];
      if (rank < 0) continue;
      const int c = (int)(((unsigned)e * cfg.magic) >> 21), k = e - c * SL;
      double* dst = G.part + ((size_t)T.q[e] * G.max_act + rank) * NV;
      const double* a = T.acc + (size_t)c * NV * SL + k;
      if constexpr (D == 3) {
        *reinterpret_cast<double2*>(dst) = make_double2(a[0], a[SL]);
        *reinterpret_cast<double2*>(dst + 2) = make_double2(a[2 * SL], a[3 * SL]);
      } else {
#pragma unroll
        for (int v = 0; v < NV; v++) dst[v] = a[v * SL];
      }
    }
    __syncwarp();
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
  double be[T], eps = 0.0, kap = 0.0;
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
    if (mtype == NLPS_MAT_DRUCKER_PRAGER) st = stress_drucker_prager<D>(mat, sp.rp, DF, be, eps, kap, tau, Wp, cep);
    else st = stress_matsuoka_nakai<D>(mat, sp.rp, DF, be, eps, kap, tau, Wp, cep);
    if (st != 0) { latch_error(err, st, P.orig[p]); return false; }
#pragma unroll
    for (int i = 0; i < T; i++) P.be_n1[(size_t)i * ld + p] = be[i];
    P.eps_n1[p] = eps;
    P.kap_n1[p] = kap;
    if (sp.rp.want_cep)
#pragma unroll
      for (int i = 0; i < D * D; i++) P.cep[(size_t)i * ld + p] = cep[i];
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
template <int D, int W, int MODE>
__global__ void __launch_bounds__(128, 4) cw_kin(const MeshDev m, const PartDev P, const GridDev G, const StepParams sp,
                                                 const CwCfg cfg, const __grid_constant__ MatTable mt, int* err,
                                                 int has_traction) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double s_tab[32];
  if (threadIdx.x < 32) s_tab[threadIdx.x] = g_exp2tab[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const CwLayout L(cfg, D, MODE);
  const WarpTile T = carve_tile(smem + (size_t)wib * L.total, L, cfg);
  constexpr int NJ = D * (D + 1) / 2, PVN = PV<D>::N;
  constexpr bool GATHER = MODE != CW_FORCE, SCATTER = MODE != CW_KIN_GATHER, FUSED = MODE == CW_KIN_FUSED;
  // per-particle record of the force sums: x (D) | lambda (D) | beta | 1/Z | G (D*D) | t (D)
  constexpr int PX = 0, PL = D, PB = 2 * D, PZ = 2 * D + 1, PG = 2 * D + 2, PT = 2 * D + 2 + D * D;
  const int SL = cfg.SL, ld = P.ld, CL = cfg.CL;
  const int nocc = *G.n_occ, nunits = (nocc + cfg.CPW - 1) / cfg.CPW;
  const int nwarps = gridDim.x * wpb, gw = blockIdx.x * wpb + wib;
  const int u_lo = (int)((long long)gw * nunits / nwarps), u_hi = (int)((long long)(gw + 1) * nunits / nwarps);
  for (int u = u_lo; u < u_hi; u++) {
    const int c0 = u * cfg.CPW, ncell = min(cfg.CPW, nocc - c0);
    cw_meta(G, nocc, P.np, c0, ncell, T, lane);
    const int t0 = T.cs[0], t1 = T.cs[ncell];
    if (u + 1 < u_hi) {  // the first rows of the next unit -> L2 while this unit computes
      cw_prefetch_rows(P.x, ld, D, t1, P.np, lane);
      cw_prefetch_rows(P.lam, ld, D, t1, P.np, lane);
      cw_prefetch_rows(P.beta, ld, 1, t1, P.np, lane);
      cw_prefetch_rows(P.zi, ld, 1, t1, P.np, lane);
      if (GATHER) cw_prefetch_rows(P.ji, ld, NJ, t1, P.np, lane);
      if (FUSED) {
        cw_prefetch_rows(P.F_n, ld, D * D, t1, P.np, lane);
        cw_prefetch_rows(P.rho, ld, 1, t1, P.np, lane);
        cw_prefetch_rows(P.vol0, ld, 1, t1, P.np, lane);
      }
      if (MODE == CW_FORCE) cw_prefetch_rows(P.gop, ld, D * D, t1, P.np, lane);
      if (lane < PPW && t1 + lane < P.np) prefetch_l2(P.clist + (size_t)(t1 + lane) * CL);
    }
    cw_stage<D, SCATTER, GATHER ? 1 : 0>(m, G, cfg, ncell, T, lane);
    if (SCATTER)
      for (int e = lane; e < ncell * SL * D; e += 32) T.acc[e] = 0.0;
    __syncwarp();
    for (int tb = t0; tb < t1; tb += PPW) {
      const int t = tb + (lane >> 2);
      const bool valid = t < t1;
      int p = 0, ci = 0, n = 0;
      if (valid) {
        p = G.plist[t];
        ci = cw_cell_of(T.cs, ncell, t);
        n = P.nnodes[p];
      }
      int stride = 0, nper = PPW, lshift = 2;
      if (FUSED) cw_pass_shape(n, cfg.NC, stride, nper, lshift);  // the weights of the gather wait in the compact cache
      const int LPP = 1 << lshift;
      for (int h = 0; h < PPW / nper; h++) {
        const int jp = lane >> lshift, subp = lane & (LPP - 1);
        const int src = (h * nper + jp) * LPP0;
        const int p_ = __shfl_sync(FULL, p, src), ci_ = __shfl_sync(FULL, ci, src), n_ = __shfl_sync(FULL, n, src);
        const bool here = __shfl_sync(FULL, (int)valid, src) != 0;
        const unsigned char* cl = P.clist + (size_t)p_ * CL;
        double* wt = T.wts + jp * stride;
        double* pvj = T.pv + jp * PVN;
        double x_[D], lam_[D], beta_ = 0.0, Zi = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) { x_[i] = 0.0; lam_[i] = 0.0; }
        if (here) {
#pragma unroll
          for (int i = 0; i < D; i++) { x_[i] = P.x[i * ld + p_]; lam_[i] = P.lam[i * ld + p_]; }
          beta_ = P.beta[p_];
          Zi = P.zi[p_];
        }
        const double* Xc_ = T.X + (size_t)ci_ * SL * D;
        if (GATHER) {
          const double* Uc_ = T.U + (size_t)ci_ * SL * D;
          double Bm[D * D];
#pragma unroll
          for (int i = 0; i < D * D; i++) Bm[i] = 0.0;
          for (int i0 = 2 * subp; i0 < n_; i0 += 2 * LPP) {
            const bool two = i0 + 1 < n_;
            const unsigned kk = *reinterpret_cast<const unsigned short*>(cl + i0);
            const int k0 = kk & 0xffu, k1 = two ? (int)(kk >> 8) : k0;
            double l0[D], l1[D], ll0 = 0.0, lx0 = 0.0, ll1 = 0.0, lx1 = 0.0;
#pragma unroll
            for (int i = 0; i < D; i++) {
              l0[i] = x_[i] - Xc_[k0 * D + i];
              l1[i] = x_[i] - Xc_[k1 * D + i];
              ll0 += l0[i] * l0[i];
              ll1 += l1[i] * l1[i];
              lx0 += l0[i] * lam_[i];
              lx1 += l1[i] * lam_[i];
            }
            const double e0 = fexp(-beta_ * ll0 + lx0, s_tab);
            const double e1 = two ? fexp(-beta_ * ll1 + lx1, s_tab) : 0.0;
            if (FUSED) {
              wt[i0] = e0;
              if (two) wt[i0 + 1] = e1;
            }
#pragma unroll
            for (int i = 0; i < D; i++) {
              const double eu0 = e0 * Uc_[k0 * D + i], eu1 = e1 * Uc_[k1 * D + i];
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
          for (int jq = 0; jq < nper; jq++) {
            const int sl = jq << lshift;
            const int nq = __shfl_sync(FULL, here ? n_ : 0, sl);
            const int cq = __shfl_sync(FULL, ci_, sl);
            const int pq_ = __shfl_sync(FULL, p_, sl);
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
            const unsigned char* clq = P.clist + (size_t)pq_ * CL;
            const double* wtq = T.wts + jq * stride;
            const double* Xq = T.X + (size_t)cq * SL * D;
            double* accq = T.acc + (size_t)cq * D * SL;
            for (int i = lane; i < nq; i += 32) {
              const int k = clq[i];
              double l[D], ll = 0.0, lx = 0.0;
#pragma unroll
              for (int d = 0; d < D; d++) {
                l[d] = xq[d] - Xq[k * D + d];
                if (MODE == CW_FORCE) { ll += l[d] * l[d]; lx += l[d] * lq_[d]; }
              }
              const double e = (MODE == CW_FORCE) ? fexp(-bq * ll + lx, s_tab) : wtq[i];
              const double N = e * zq;
#pragma unroll
              for (int d = 0; d < D; d++) {
                double gl = tq[d];
#pragma unroll
                for (int kk = 0; kk < D; kk++) gl += Gq[d * D + kk] * l[kk];
                accq[d * SL + k] += N * gl;
              }
            }
            __syncwarp();
          }
        }
        __syncwarp();
      }
    }
    if (SCATTER) {
      for (int e = lane; e < ncell * SL; e += 32) {
        const int rank = T.rank[e];
        if (rank < 0) continue;
        const int c = (int)(((unsigned)e * cfg.magic) >> 21), k = e - c * SL;
        double* dst = G.part + ((size_t)T.q[e] * G.max_act + rank) * D;
        const double* a = T.acc + (size_t)c * D * SL + k;
#pragma unroll
        for (int v = 0; v < D; v++) dst[v] = a[v * SL];
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// K4: G2P + corrector (U-Verlet.c:963-1084): a_p = sum N_A a_A, DU_p = sum N_A DU_A, v += gamma dt a, x += DU,
// dis += DU.  The n+1 -> n roll of F, J, b_e, kappa, EPS is a pointer swap on the host side of the engine.
template <int D, int W>
__global__ void __launch_bounds__(128, 4) cw_g2p(const MeshDev m, const PartDev P, const GridDev G, const StepParams sp,
                                                 const CwCfg cfg) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double s_tab[32];
  if (threadIdx.x < 32) s_tab[threadIdx.x] = g_exp2tab[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const CwLayout L(cfg, D, CW_G2P);
  const WarpTile T = carve_tile(smem + (size_t)wib * L.total, L, cfg);
  const int SL = cfg.SL, ld = P.ld, CL = cfg.CL;
  const int nocc = *G.n_occ, nunits = (nocc + cfg.CPW - 1) / cfg.CPW;
  const int nwarps = gridDim.x * wpb, gw = blockIdx.x * wpb + wib;
  const int u_lo = (int)((long long)gw * nunits / nwarps), u_hi = (int)((long long)(gw + 1) * nunits / nwarps);
  for (int u = u_lo; u < u_hi; u++) {
    const int c0 = u * cfg.CPW, ncell = min(cfg.CPW, nocc - c0);
    cw_meta(G, nocc, P.np, c0, ncell, T, lane);
    const int t0 = T.cs[0], t1 = T.cs[ncell];
    if (u + 1 < u_hi) {
      cw_prefetch_rows(P.x, ld, D, t1, P.np, lane);
      cw_prefetch_rows(P.lam, ld, D, t1, P.np, lane);
      cw_prefetch_rows(P.vel, ld, D, t1, P.np, lane);
      cw_prefetch_rows(P.dis, ld, D, t1, P.np, lane);
      cw_prefetch_rows(P.beta, ld, 1, t1, P.np, lane);
      cw_prefetch_rows(P.zi, ld, 1, t1, P.np, lane);
      if (lane < PPW && t1 + lane < P.np) prefetch_l2(P.clist + (size_t)(t1 + lane) * CL);
    }
    cw_stage<D, false, 2>(m, G, cfg, ncell, T, lane);
    __syncwarp();
    for (int tb = t0; tb < t1; tb += PPW) {
      // 4 lanes per particle, lane `sub` takes the neighbour pairs (2 sub, 2 sub + 1), + 8, ...
      const int sub = lane & (LPP0 - 1);
      const int t = tb + (lane >> 2);
      const bool valid = t < t1;
      int p = 0, ci = 0, n = 0;
      double x_[D], lam_[D], a[D], du[D], beta_ = 0.0, Zi = 0.0, vel_s = 0.0, dis_s = 0.0;
#pragma unroll
      for (int i = 0; i < D; i++) { x_[i] = 0.0; lam_[i] = 0.0; a[i] = 0.0; du[i] = 0.0; }
      if (valid) {
        p = G.plist[t];
        ci = cw_cell_of(T.cs, ncell, t);
        n = P.nnodes[p];
#pragma unroll
        for (int i = 0; i < D; i++) { x_[i] = P.x[i * ld + p]; lam_[i] = P.lam[i * ld + p]; }
        beta_ = P.beta[p];
        Zi = P.zi[p];
        // component `sub` of the particle is updated by lane `sub`: its old values are requested now, used after the loop
        if (sub < D) { vel_s = P.vel[sub * ld + p]; dis_s = P.dis[sub * ld + p]; }
      }
      const unsigned char* cl = P.clist + (size_t)p * CL;
      const double* Xc_ = T.X + (size_t)ci * SL * D;
      const double* Uc_ = T.U + (size_t)ci * SL * D;
      const double* Ac_ = T.A + (size_t)ci * SL * D;
      for (int i0 = 2 * sub; i0 < n; i0 += 2 * LPP0) {
        const bool two = i0 + 1 < n;
        const unsigned kk = *reinterpret_cast<const unsigned short*>(cl + i0);
        const int k0 = kk & 0xffu, k1 = two ? (int)(kk >> 8) : k0;
        double ll0 = 0.0, lx0 = 0.0, ll1 = 0.0, lx1 = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) {
          const double l0 = x_[i] - Xc_[k0 * D + i], l1 = x_[i] - Xc_[k1 * D + i];
          ll0 += l0 * l0;
          ll1 += l1 * l1;
          lx0 += l0 * lam_[i];
          lx1 += l1 * lam_[i];
        }
        const double e0 = fexp(-beta_ * ll0 + lx0, s_tab);
        const double e1 = two ? fexp(-beta_ * ll1 + lx1, s_tab) : 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) {
          a[i] += e0 * Ac_[k0 * D + i];
          du[i] += e0 * Uc_[k0 * D + i];
          a[i] += e1 * Ac_[k1 * D + i];
          du[i] += e1 * Uc_[k1 * D + i];
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
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
template <class K>
int cw_prepare(K kernel, int id, const CwLaunch& L, CwState& st, int D, int sm_count, int max_smem_optin) {
  if (st.ready[id]) return 0;
  const size_t per_warp = CwLayout(L.cfg, D, id).total;
  st.smem[id] = per_warp * L.cfg.warps;
  if (st.smem[id] > (size_t)max_smem_optin) {
    fprintf(stderr, "nlps_b200: 2-ring too large for the warp tiles (%zu bytes of shared memory per block)\n", st.smem[id]);
    return 1;
  }
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

size_t cw_smem_bytes(int D, int kernel, const CwCfg& cfg) { return CwLayout(cfg, D, kernel).total * cfg.warps; }

int cw_launch_lme_p2g(int D, int W, const CwLaunch& L, CwState& st, int sm_count, int max_smem_optin, int do_predictor) {
#define GO_(d, w)                                                                                                     \
  {                                                                                                                   \
    auto kf = cw_lme_p2g<d, w>;                                                                                       \
    if (cw_prepare(kf, CW_LME_P2G, L, st, d, sm_count, max_smem_optin)) return 1;                                      \
    kf<<<std::min(L.max_blocks, st.grid[CW_LME_P2G]), L.cfg.warps * 32, st.smem[CW_LME_P2G], L.stream>>>(              \
        L.m, L.P, L.G, L.sp, L.cfg, L.err, do_predictor);                                                             \
    return 0;                                                                                                         \
  }
  if (D == 2) { if (W == 1) GO_(2, 1) if (W == 2) GO_(2, 2) }
  else { if (W == 4) GO_(3, 4) if (W == 8) GO_(3, 8) }
#undef GO_
  return 1;
}

int cw_launch_kin(int D, int W, int mode, const CwLaunch& L, CwState& st, int sm_count, int max_smem_optin,
                  const MatTable& mt, int has_traction) {
#define GO_(d, w, md)                                                                                                 \
  {                                                                                                                   \
    auto kf = cw_kin<d, w, md>;                                                                                       \
    if (cw_prepare(kf, md, L, st, d, sm_count, max_smem_optin)) return 1;                                              \
    kf<<<std::min(L.max_blocks, st.grid[md]), L.cfg.warps * 32, st.smem[md], L.stream>>>(L.m, L.P, L.G, L.sp, L.cfg,   \
                                                                                         mt, L.err, has_traction);    \
    return 0;                                                                                                         \
  }
#define MODE_(d, w)                                                                                                   \
  {                                                                                                                   \
    if (mode == CW_KIN_FUSED) GO_(d, w, CW_KIN_FUSED)                                                                  \
    if (mode == CW_KIN_GATHER) GO_(d, w, CW_KIN_GATHER)                                                                \
    if (mode == CW_FORCE) GO_(d, w, CW_FORCE)                                                                          \
  }
  if (D == 2) { if (W == 1) MODE_(2, 1) if (W == 2) MODE_(2, 2) }
  else { if (W == 4) MODE_(3, 4) if (W == 8) MODE_(3, 8) }
#undef MODE_
#undef GO_
  return 1;
}

int cw_launch_stress(int D, const CwLaunch& L, const MatTable& mt, int uniform_mat, int has_traction) {
  (void)has_traction;
  const int np = L.P.np;
  if (np <= 0) return 0;
  const int grid = (np + 127) / 128;
#define GO_(d, mat) { cw_stress<d, mat><<<grid, 128, 0, L.stream>>>(L.P, L.sp, mt, L.err); return 0; }
#define MAT_(d) switch (uniform_mat) { case 0: GO_(d, 0) case 1: GO_(d, 1) case 2: GO_(d, 2) default: GO_(d, -1) }
  if (D == 2) MAT_(2) else MAT_(3)
#undef MAT_
#undef GO_
  return 1;
}

int cw_launch_g2p(int D, int W, const CwLaunch& L, CwState& st, int sm_count, int max_smem_optin) {
#define GO_(d, w)                                                                                                     \
  {                                                                                                                   \
    auto kf = cw_g2p<d, w>;                                                                                           \
    if (cw_prepare(kf, CW_G2P, L, st, d, sm_count, max_smem_optin)) return 1;                                          \
    kf<<<std::min(L.max_blocks, st.grid[CW_G2P]), L.cfg.warps * 32, st.smem[CW_G2P], L.stream>>>(L.m, L.P, L.G, L.sp,  \
                                                                                                 L.cfg);              \
    return 0;                                                                                                         \
  }
  if (D == 2) { if (W == 1) GO_(2, 1) if (W == 2) GO_(2, 2) }
  else { if (W == 4) GO_(3, 4) if (W == 8) GO_(3, 8) }
#undef GO_
  return 1;
}
