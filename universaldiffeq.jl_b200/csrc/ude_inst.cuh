// Helper for the ude_inst_*.cu translation units: instantiate ude_seg_kernel for one (RHS, G, HPL) combination
// (both solvers, with and without the reverse sweep) and register the four kernels.
#pragma once
#include "ude_kernel.cuh"
#include "ude_registry.h"

namespace ude {
template <class R, int G, int HPL, int MINB, int SOLVER, int LK, bool WS>
void register_one(const char* name, int priority) {
  using GG = Geo<R, G, HPL>;
  const int wsd = WS ? G * GG::WREC2 * 2 : 0;
  kernel_registry().push_back(KernelEntry{R::KIND, R::D, R::NX, R::DIN, R::DOUT, G, HPL, SOLVER, LK, 1, GG::ROWS, wsd, LK != LK_DM, priority,
                                          (const void*)ude_seg_kernel<R, G, HPL, SOLVER, LK, true, MINB, WS>, name});
  kernel_registry().push_back(KernelEntry{R::KIND, R::D, R::NX, R::DIN, R::DOUT, G, HPL, SOLVER, LK, 0, GG::ROWS, wsd, 0, priority,
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
}  // namespace ude
