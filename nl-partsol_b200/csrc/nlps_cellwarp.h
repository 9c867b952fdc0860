// nlps_cellwarp.h -- host-side interface of the warp-per-cell kernels (nlps_cellwarp.cu), called by the stage
// functions of nlps_engine.cu.  Plain C++ (no CUDA types beyond cudaStream_t), one translation unit each side.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "nlps_types.cuh"

// Geometry of a warp's work unit (3D).  A warp owns one occupied cell at a time (a cell = all particles with the same
// closest node I0) and walks its particles in chunks of 8; a particle's neighbour loop is split over 4 lanes (8..32 in
// the fused kinematics kernel when a list is longer than the compact cache of its weights, NC entries per particle).
struct CwCfg {
  int SL;          // longest 2-ring row
  int NC;          // compact weight cache entries per particle (8 particles per chunk share 8 * NC entries)
  int CL;          // bytes per particle of P.clist (SL rounded up to a multiple of 4)
  int W;           // mask words per particle (4 or 8)
  int warps;       // warps per block
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

size_t cw_smem_bytes(int kernel, const CwCfg& cfg);
// every launcher returns 0 or 1 (bad configuration); kernel failures surface through the engine's error latch
// ndim must be 3 (2D decks run the cell-group kernels of nlps_engine.cu); W = mask words (4 or 8)
int cw_launch_lme_p2g(int ndim, int W, const CwLaunch& L, CwState& st, int sm_count, int max_smem_optin, int do_predictor);
// mode: CW_KIN_FUSED (Neo-Hookean clouds: kinematics + stress + force sums in one kernel), CW_KIN_GATHER (kinematics
// only, writes DF), CW_FORCE (force sums from the force operator written by cw_launch_stress)
int cw_launch_kin(int ndim, int mode, const CwLaunch& L, CwState& st, int sm_count, int max_smem_optin, const MatTable& mt,
                  int has_traction);
// thread per particle: F_n1, J, rho, stress, history, force operator from DF (plastic and mixed clouds)
int cw_launch_stress(int ndim, const CwLaunch& L, const MatTable& mt, int uniform_mat);
int cw_launch_g2p(int ndim, const CwLaunch& L, CwState& st, int sm_count, int max_smem_optin);
