// nlps_types.cuh -- device views and inline device helpers shared by the translation units of the engine
// (nlps_engine.cu: host orchestration + node / scan / cell-block kernels; nlps_cellwarp.cu: warp-per-cell kernels).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nlps_b200.h"
#include "nlps_device.cuh"

static const int MAX_MATERIALS = 8;
static const int MAX_MASK_WORDS = 8;  // 2-ring up to 256 nodes
// material table of ONE engine: travels by value as a kernel parameter (constant bank), so that engines with different
// decks can live in one process
struct MatTable { MatParams m[MAX_MATERIALS]; };

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
// D doubles from shared memory: one 16-byte load in 2D (the staged node arrays are 16-byte aligned, stride 2)
template <int D>
__device__ __forceinline__ void ldsvec(const double* p, double* out) {
  if constexpr (D == 2) {
    const double2 a = *reinterpret_cast<const double2*>(p);
    out[0] = a.x; out[1] = a.y;
  } else {
#pragma unroll
    for (int i = 0; i < D; i++) out[i] = p[i];
  }
}
struct MeshDev {
  int nn;
  const double* X;  // nn x NS<D>::X (row-major, padded)
  const int *r1p, *r1i, *r2p, *r2i;
  const int *r1tp, *r1ti, *r2tp, *r2ti;  // transposed adjacency (who lists me)
  const unsigned char* r2q;              // r2q[r2p[B]+s] = position of B inside the r2t row of node r2i[r2p[B]+s]
  const double* h_avg;
  const double* sst;  // per node: s* of beta = gamma / h_avg^2 (what a particle of that cell tests its neighbours with next step)
  // cell-major cell sums: the record of ring slot s of cell B sits at position r2pos[r2p[B]+s] of the cell's run, and
  // r2ts[r2tp[A]+q] is the position of node A's record in the run of the cell r2ti[r2tp[A]+q].  3D: position = slot.  2D
  // (a thread per (cell, slot) pair writes, any order is free): position = rank of the slot's node id in the ring, so that
  // the records x-consecutive nodes read from a cell are adjacent -- one 128-byte line serves four nodes of a warp.
  const unsigned char* r2ts;
  const unsigned char* r2pos;
};

struct PartDev {
  int np;  // particles held by this engine (changes when particles migrate between slabs)
  int ld;  // leading dimension of the SoA arrays = capacity (np <= ld)
  // SoA, component-major: f[c*ld + p]; p is the PHYSICAL slot (cell-sorted every few steps),
  // orig[p] the caller's (global) particle id and inv[] its inverse (-1: not held by this slab).
  double *x, *dis, *ddis, *vel, *acc, *lam;
  double *beta, *mass, *vol0, *rho, *W;
  double *J_n, *J_n1, *eps_n, *eps_n1, *kap_n, *kap_n1;
  double *F_n, *F_n1, *DF, *be_n, *be_n1, *stress, *cep;
  double *Fs4, *DFs4;  // 2D slot 4 of F / DF (never touched by the kinematics, Appendix B)
  double* trac;        // D x np: Neumann traction * A0 of the current step (allocated only with loads)
  double* area0;       // Phi.Area_0 (3D decks with Neumann loads only, else nullptr)
  double* back;        // Phi.Back_stress, 3 principal components (clouds with a Von-Mises material only, else nullptr)
  // shape-function data of the current step, written by the LME kernel and read by the kinematics / force / G2P
  // kernels (same x_p, lambda, beta within a step): 1 / Z and the inverse Hessian J^-1 (symmetric, D(D+1)/2 entries)
  double *zi, *ji;
  double* sstar;       // s* of the particle's beta: "s <= s*" is the reference's "sqrt(s) <= Ra" (LME.c:1052,1074)
  unsigned char* clist;  // CL bytes per particle: the neighbour list of the step as ascending 2-ring slot ids
  double* gop;         // D*D x np: force operator V0 tau DF^-T J^-1 of the current step (stress kernel -> force kernel)
  int *I0, *nnodes, *matidx, *orig, *inv;
  uint32_t* mask;  // W words, word-major: mask[w*np + p]
};

struct GridDev {
  double *M, *F;  // M: nn ; F: nn x D (row-major)
  double* MOM;    // nn x D: sum m N DU_p before the division by M (kept for the slab halo sums)
  unsigned char* rocc;  // cell occupied by particles of a NEIGHBOUR slab (multi-GPU), zero otherwise
  unsigned char* band;  // node lies in the halo band of one of the slab's cuts (its sums wait for the exchange), or nullptr
  double* UA;     // per node [dU (NS) | A (NS)]: the two nodal fields the G2P gathers read, one record
  unsigned char *active, *fixed;
  int *cnt, *cursor, *cell_start, *plist, *act_list, *n_active;
  int *occ_list, *n_occ, *act_pos, *occ_pos;
  int4* occ_meta;   // per occupied cell (in node order): {node B, first particle slot, 2-ring base, 2-ring length | particles << 9}
  int* arank;       // rank of a node among the active nodes, -1 when inactive
  uint32_t* occm;   // per active rank: transposed-2-ring slots whose cell is occupied (w2t words, word-major)
  ulonglong2 *packed, *scan_blk;
  // per block of 256 node ids: holds an occupied cell (set by the search) / lies in the 2-ring of one (this step,
  // previous step) / must be processed by the node kernels this step (= dirty now or last step: leaving blocks are
  // visited once more so that their flags return to zero)
  unsigned char *occ_blk, *dirty_cur, *dirty_prev, *live;
  double* part;     // cell partial sums.  Slot-major (2D kernels): part[(q * max_act + rank) * NV + v], coalesced for the node
                    // kernels that read it.  Cell-major (3D warp-per-cell kernels, cm_sl > 0): part[(cell rank * cm_sl + slot) * 4 + v]:
                    // a cell's sums leave the SM as ONE contiguous run -- the scattered 8-byte stores of the slot-major layout
                    // were a third of cw_kin (profiles/ab/r02_probe_phases.txt); the node kernels gather 32-byte records instead
  int cap, max_act, w2t;
  int cm_sl;        // 0: slot-major part[]; > 0: cell-major with this many slots per cell
  uint32_t* cum;    // cell-major: per occupied cell, the slots that carry a non-zero record (cm_w words per cell)
  int cm_w;
};

struct StepParams {
  double dt, gamma_lme, neg_log_tol, tol_wrapper, thickness;
  int max_iter_lme, nsteps, step, update_I0, W;
  ReturnMapParams rp;
  // implicit scheme (U-Newmark-beta.c): project `proj` (D x ld SoA) instead of D_dis, keep the neighbour lists and
  // beta of the search already done this step, leave the density alone in the kinematics
  const double* proj;
  int reuse_lists, implicit;
};

// slab view of the kernels: ownership interval [own_lo, own_hi) of closest-node coordinates along `axis`
// and the wider interval [lo, hi] a particle may roam between two migrations (halo band minus 3.5 cells)
struct SlabDev { int on, axis; double lo, hi, own_lo, own_hi; };

// One thread block works on C consecutive OCCUPIED cells (a cell = all particles with the same closest
// node I0) = one contiguous run of the cell-sorted particle order.  SL = longest 2-ring row, PCAP =
// particles whose per-particle scratch fits in shared memory at once (longer runs go in chunks).
struct BlockCfg { int C, SL, PCAP, threads; unsigned magic; int NCA, NCB; int cellfast; };  // cellfast: bit 0 / 1 = cell-fastest pair order in the uncached cell phase of k_lme_p2g / k_kin_force  // NCA / NCB: compact weight cache entries per particle in k_lme_p2g / k_kin_force (0 = none)  // magic = ceil(2^21 / SL): e / SL == (e * magic) >> 21 for e < C*SL (checked at create)


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

// aLME (Nodes/aLME.c:382-434): the weight exponent is -l^T B l + lambda.l with the particle's metric B instead of
// -beta |l|^2 + lambda.l.  The 2D metric is kept as (B00, B01 + B10, B11); `on` is uniform over a launch.
struct MetricB { double b00, bs, b11; int on; };
// the per-particle aLME arrays (2D: 4 columns each, component-major with the leading dimension of PartDev): the
// thermalisation metric (Particle.Beta) and the cut-off ellipsoid of the neighbour test; both nullptr with LME.  Kept out
// of PartDev and handed to the 2D kernels as an argument of its own, so that the 3D warp-per-cell kernels keep their
// parameter layout (and their machine code: the ncu captures under profiles/ are keyed by it)
struct AlmeDev { double *bten, *cten; };
template <int D>
__device__ __forceinline__ MetricB metric_load(const AlmeDev& al, int ld, int p) {
  MetricB M;
  M.on = 0; M.b00 = M.bs = M.b11 = 0.0;
  if (D == 2 && al.bten) {
    M.on = 1;
    M.b00 = al.bten[p];
    M.bs = al.bten[(size_t)ld + p] + al.bten[(size_t)2 * ld + p];
    M.b11 = al.bten[(size_t)3 * ld + p];
  }
  return M;
}
// beta |l|^2 (LME.c:700-737) or l^T B l (aLME.c:382-405)
template <int D>
__device__ __forceinline__ double metric_q(const MetricB& M, double beta, double ll, const double* l) {
  if (D == 2 && M.on) return M.b00 * (l[0] * l[0]) + M.bs * (l[0] * l[1]) + M.b11 * (l[1] * l[1]);
  return beta * ll;
}
// M <- DF^-T M DF^-1 (update_beta__aLME__ / update_cut_off_ellipsoid__aLME__, aLME.c:693-809), 2 x 2 row-major, the four
// terms of every entry in the reference's order
__device__ __forceinline__ void alme_push_forward(const double* Fi, double* M) {
  double U[4];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 2; j++) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 2; k++)
#pragma unroll
        for (int l = 0; l < 2; l++) s += Fi[k * 2 + i] * M[k * 2 + l] * Fi[l * 2 + j];
      U[i * 2 + j] = s;
    }
#pragma unroll
  for (int i = 0; i < 4; i++) M[i] = U[i];
}
// generalised_Euclidean_distance__MatrixLib__ (MatrixOp.c:895-920) with the reference's rounding sequence (products and
// sums rounded separately): sqrt(l^T C l), compared with 1 by tributary__aLME__ (aLME.c:811-872)
__device__ __forceinline__ double alme_distance(const double* C, const double* l) {
  double q = 0.0;
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const double Cl = __dadd_rn(__dmul_rn(C[i * 2 + 0], l[0]), __dmul_rn(C[i * 2 + 1], l[1]));
    q = __dadd_rn(q, __dmul_rn(l[i], Cl));
  }
  return __dsqrt_rn(q);
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

// exp() for the LME weights: exp(x) = 2^(k/32) * exp(r), k = rint(32 x / ln 2), |r| <= ln2/64, table of
// 2^(j/32) (shared memory) times a degree-6 Taylor polynomial: 11 fp64 operations instead of libdevice's 17
// plus constant moves, max relative error 1.94e-16 (0.9 ulp; checked against mpmath over [-700, 700]).
// The kernels issue two of these chains per loop iteration (for_neighbour_pairs): the fp64 pipe, not the
// latency of one dependent chain, then bounds the shape-function loops.
static __device__ const double g_exp2tab[32] = {
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237, 1.0905077326652577, 1.1143867425958924,
    1.1387886347566916, 1.1637248587775775, 1.189207115002721, 1.215247359980469, 1.241857812073484,
    1.2690509571917332, 1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,
    1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228, 1.5422108254079407,
    1.5759808451078865, 1.6104903319492543, 1.645755478153965, 1.681792830507429, 1.718619298122478,
    1.7562521603732995, 1.7947090750031072, 1.8340080864093424, 1.8741676341103, 1.9152065613971474,
    1.9571441241754002};
// constants in the constant bank: a DFMA takes them as a direct operand (an immediate would cost two moves each)
static __constant__ double c_fexp[8] = {46.16624130844683,        // 32 / ln 2
                                 -0.02166084938653512,     // -ln2/32, high part (21 trailing zero bits: exact product)
                                 -5.9631716539705866e-12,  // -ln2/32, low part
                                 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.0};
__device__ __forceinline__ double fexp(double x, const double* tab) {
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: the low word of x*c + MAGIC is rint(x*c)
  const double kd = __fma_rn(x, c_fexp[0], MAGIC);
  const int k = __double2loint(kd);
  const double kf = kd - MAGIC;
  double r = __fma_rn(kf, c_fexp[1], x);
  r = __fma_rn(kf, c_fexp[2], r);
  double q = __fma_rn(r, c_fexp[3], c_fexp[4]);
  q = __fma_rn(r, q, c_fexp[5]);
  q = __fma_rn(r, q, c_fexp[6]);
  q = __fma_rn(r, q, 0.5);
  q = __fma_rn(r, q, 1.0);
  // 2^(k/32) = table entry with the exponent shifted (clamped: |x| > 708 saturates instead of wrapping;
  // a NaN argument still gives NaN through r)
  const int m = max(-1021, min(1022, k >> 5));
  const double T = tab[k & 31];
  const double Ts = __hiloint2double(__double2hiint(T) + (m << 20), __double2loint(T));
  return __fma_rn(Ts, r * q, Ts);
}
// exp() without a table (warp-per-cell kernels, whose busiest unit is the shared-memory / L1 data pipe: the table
// look-up of fexp() was a fifth of their shared-memory wavefronts): exp(x) = 2^k exp(r), k = rint(x / ln 2),
// |r| <= ln2 / 2, degree-13 Taylor polynomial in Horner form: 17 fp64 operations, max relative error 1.4e-16 (checked
// against mpmath over [-700, 700]).
__constant__ static double c_fexp2[16] = {1.4426950408889634,        // 1 / ln 2
                                         -0.693147180369123816490,  // -ln2, high part (21 trailing zero bits: exact product)
                                         -1.90821492927058770002e-10,  // -ln2, low part
                                         1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0,
                                         1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0,
                                         1.0 / 6.0, 0.5, 0.0};
__device__ __forceinline__ double fexp_poly(double x) {
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: the low word of x*c + MAGIC is rint(x*c)
  const double kd = __fma_rn(x, c_fexp2[0], MAGIC);
  const int k = __double2loint(kd);
  const double kf = kd - MAGIC;
  double r = __fma_rn(kf, c_fexp2[1], x);
  r = __fma_rn(kf, c_fexp2[2], r);
  // exp(r) - 1 - r = r^2 (c2 + c3 r + ... + c13 r^11), Estrin's scheme: the dependent chain is 6 operations deep
  // instead of the 13 of Horner's (these kernels run 4 warps per scheduler: the chain length is what they wait on)
  const double r2 = r * r, r4 = r2 * r2;
  const double a0 = __fma_rn(r, c_fexp2[13], c_fexp2[14]);  // c2 + c3 r
  const double a1 = __fma_rn(r, c_fexp2[11], c_fexp2[12]);  // c4 + c5 r
  const double a2 = __fma_rn(r, c_fexp2[9], c_fexp2[10]);   // c6 + c7 r
  const double a3 = __fma_rn(r, c_fexp2[7], c_fexp2[8]);    // c8 + c9 r
  const double a4 = __fma_rn(r, c_fexp2[5], c_fexp2[6]);    // c10 + c11 r
  const double a5 = __fma_rn(r, c_fexp2[3], c_fexp2[4]);    // c12 + c13 r
  const double b0 = __fma_rn(r2, a1, a0), b1 = __fma_rn(r2, a3, a2), b2 = __fma_rn(r2, a5, a4);
  const double q = __fma_rn(r4, __fma_rn(r4, b2, b1), b0);
  const double t = __fma_rn(r2, q, r);
  // 2^k by the exponent field (clamped: |x| > 708 saturates instead of wrapping; a NaN argument gives NaN through t)
  const int m = max(-1022, min(1023, k));
  const double s = __hiloint2double((m + 1023) << 20, 0);
  return __fma_rn(s, t, s);
}
// visit the set bits of a neighbour mask two at a time; `two` is false for the odd one out (k1 == k0)
template <int W, class F>
__device__ __forceinline__ void for_neighbour_pairs(const uint32_t* mk, F&& f) {
#pragma unroll
  for (int w = 0; w < W; w++) {
    uint32_t mm = mk[w];
    while (mm) {
      const int b0 = __ffs(mm) - 1;
      mm &= mm - 1;
      const bool two = mm != 0u;
      const int b1 = two ? __ffs(mm) - 1 : b0;
      mm &= mm - 1;
      f(w * 32 + b0, w * 32 + b1, two);
    }
  }
}

// 2D: walk ALL slots of the cell's 2-ring in order, two at a time, with weight 0 for the slots that are not
// neighbours: the lanes of a warp that sit in the same cell then read the same shared-memory words in the same
// instruction (broadcast: one wavefront instead of one per lane -- the LSU data pipe is the busiest unit of these
// kernels, profiles/r01_ncu_full_c2_v5.txt) and the bit scanning disappears; the price is 25 instead of ~22 exp
// evaluations, which the fp64 pipe (16-25 % busy) absorbs.  3D (125 slots, ~40 neighbours) keeps the bit iteration.
template <int D, int W, class F>
__device__ __forceinline__ void for_slots(const uint32_t* mk, int len, F&& f) {
  if constexpr (D == 2 && W == 1) {
    const uint32_t m = mk[0];
    for (int k = 0; k < len; k += 2) {
      const bool has1 = k + 1 < len;
      const double w0 = ((m >> k) & 1u) ? 1.0 : 0.0;
      const double w1 = (has1 && ((m >> (k + 1)) & 1u)) ? 1.0 : 0.0;
      f(k, has1 ? k + 1 : k, w0, w1);
    }
  } else {
    for_neighbour_pairs<W>(mk, [&](int k0, int k1, bool two) { f(k0, k1, 1.0, two ? 1.0 : 0.0); });
  }
}
template <int D, int W> struct DenseSlots { static constexpr bool value = (D == 2 && W == 1); };
