// Helper for the ude_inst_*.cu translation units: instantiate ude_seg_kernel for one (RHS, G, HPL) combination
// (both solvers, with and without the reverse sweep) and register the four kernels.
#pragma once
#include "ude_kernel.cuh"
#include "ude_kernel_mma.cuh"
#include "ude_registry.h"

namespace ude {
template <class R, int G, int HPL, int MINB, int SOLVER, int LK, bool WS>
void register_one(const char* name, int priority) {
  using GG = Geo<R, G, HPL>;
  const int wsd = WS ? G * GG::WREC2 * 2 : 0;
  const int ring = (LK != LK_DM) ? 2 * Tab<SOLVER>::S * GG::ROWS * 32 : 0;
  kernel_registry().push_back(KernelEntry{R::KIND, R::D, R::NX, R::DIN, R::DOUT, G, HPL, SOLVER, LK, 1, GG::ROWS, wsd, 0, wsd + ring, 0, 32 / G, 1, priority,
                                          (const void*)ude_seg_kernel<R, G, HPL, SOLVER, LK, true, MINB, WS>, name});
  kernel_registry().push_back(KernelEntry{R::KIND, R::D, R::NX, R::DIN, R::DOUT, G, HPL, SOLVER, LK, 0, GG::ROWS, wsd, 0, wsd, 0, 32 / G, 1, priority,
                                          (const void*)ude_seg_kernel<R, G, HPL, SOLVER, LK, false, MINB, WS>, name});
}
template <class R, int G, int HPL, int MINB, bool WS = false>
void register_combo(const char* name, int priority) {
  register_one<R, G, HPL, MINB, 0, LK_STATE, WS>(name, priority);
  register_one<R, G, HPL, MINB, 0, LK_TRAJ, WS>(name, priority);
  register_one<R, G, HPL, MINB, 0, LK_DM, WS>(name, priority);
  register_one<R, G, HPL, MINB, 1, LK_STATE, WS>(name, priority);
  register_one<R, G, HPL, MINB, 1, LK_TRAJ, WS>(name, priority);
  register_one<R, G, HPL, MINB, 1, LK_DM, WS>(name, priority);
}

// tensor-core (FP64 DMMA) family: NODE right-hand sides, 8 segments per task, HW warps per task each owning HT hidden
// tiles (hidden width padded to 8*HT*HW)
template <int D, int NX, int HT, int HW, int MINB, int SOLVER, int LK, int TB = 4>
void register_mma_one(const char* name, int priority) {
  using MG = MmaGeo<D, NX, HT, HW, TB>;
  // registry fields G / HPL are reused as "segments per warp-task = 32/G" and capacity G*HPL >= nn_hidden
  const int G = 4, HPLcap = 2 * HT * HW;
  kernel_registry().push_back(KernelEntry{UDE_RHS_NODE, D, NX, D + NX, D, G, HPLcap, SOLVER, LK, 1, MG::REC / 32, 0,
                                          MmaSmem<MG, LK, true>::BLOCK, MmaSmem<MG, LK, true>::WARP, 1, 8, HW, priority,
                                          (const void*)ude_mma_kernel<D, NX, HT, HW, SOLVER, LK, true, MINB, TB>, name});
  kernel_registry().push_back(KernelEntry{UDE_RHS_NODE, D, NX, D + NX, D, G, HPLcap, SOLVER, LK, 0, MG::REC / 32, 0,
                                          MmaSmem<MG, LK, false>::BLOCK, MmaSmem<MG, LK, false>::WARP, 1, 8, HW, priority,
                                          (const void*)ude_mma_kernel<D, NX, HT, HW, SOLVER, LK, false, MINB, TB>, name});
}
template <int D, int NX, int HT, int MINB, int HW = 1, int TB = 4>
void register_mma_combo(const char* name, int priority) {
  register_mma_one<D, NX, HT, HW, MINB, 0, LK_STATE, TB>(name, priority);
  register_mma_one<D, NX, HT, HW, MINB, 0, LK_TRAJ, TB>(name, priority);
  register_mma_one<D, NX, HT, HW, MINB, 0, LK_DM, TB>(name, priority);
  register_mma_one<D, NX, HT, HW, MINB, 1, LK_STATE, TB>(name, priority);
  register_mma_one<D, NX, HT, HW, MINB, 1, LK_TRAJ, TB>(name, priority);
  register_mma_one<D, NX, HT, HW, MINB, 1, LK_DM, TB>(name, priority);
}
}  // namespace ude
