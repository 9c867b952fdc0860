// nlps_cellwarp.h -- host-side interface of the warp-per-cell kernels (nlps_cellwarp.cu), called by the stage
// functions of nlps_engine.cu.  Plain C++ (no CUDA types beyond cudaStream_t), one translation unit each side.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "nlps_types.cuh"

// Geometry of a warp's work unit.  A warp owns CPW consecutive occupied cells (a cell = all particles with the same
// closest node I0) and walks their particles in chunks of 8; a particle's neighbour loop is split over 4 (or, for long
// lists, 8..32) lanes.  NC = entries per particle of the compact neighbour cache (slot ids + weights) in shared memory.
struct CwCfg {
  int SL;          // longest 2-ring row
  int CPW;         // cells per warp unit
  int NC;          // compact cache entries per particle (8 particles per chunk share 8 * NC entries)
  int CL;          // bytes per particle of P.clist (SL rounded up to a multiple of 4)
  int warps;       // warps per block
  unsigned magic;  // ceil(2^21 / SL): e / SL == (e * magic) >> 21 for e < CPW * SL
};

enum CwKernel { CW_LME_P2G = 0, CW_KIN_FUSED = 1, CW_KIN_GATHER = 2, CW_FORCE = 3, CW_G2P = 4, CW_KERNELS = 5 };

struct CwLaunch {
  MeshDev m;
  PartDev P;
  GridDev G;
  StepParams sp;
  CwCfg cfg;
  int* err;
  cudaStream_t stream;
  int max_blocks;   // upper bound from the problem size (units / warps per block)
};

// per-engine launch state of one kernel instantiation: resident blocks and dynamic shared memory, resolved once
struct CwState {
  int grid[CW_KERNELS] = {0, 0, 0, 0, 0};
  size_t smem[CW_KERNELS] = {0, 0, 0, 0, 0};
  int ready[CW_KERNELS] = {0, 0, 0, 0, 0};
};

size_t cw_smem_bytes(int D, int kernel, const CwCfg& cfg);
// every launcher returns 0 or 1 (bad configuration); kernel failures surface through the engine's error latch
int cw_launch_lme_p2g(int D, int W, const CwLaunch& L, CwState& st, int sm_count, int max_smem_optin, int do_predictor);
// mode: CW_KIN_FUSED (Neo-Hookean clouds: kinematics + stress + force sums in one kernel), CW_KIN_GATHER (kinematics
// only, writes DF), CW_FORCE (force sums from the force operator written by cw_launch_stress)
int cw_launch_kin(int D, int W, int mode, const CwLaunch& L, CwState& st, int sm_count, int max_smem_optin,
                  const MatTable& mt, int has_traction);
// thread per particle: F_n1, J, rho, stress, history, force operator from DF (plastic and mixed clouds)
int cw_launch_stress(int D, const CwLaunch& L, const MatTable& mt, int uniform_mat, int has_traction);
int cw_launch_g2p(int D, int W, const CwLaunch& L, CwState& st, int sm_count, int max_smem_optin);
