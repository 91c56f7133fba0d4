// Helper for the ude_inst_*.cu translation units: instantiate ude_seg_kernel for one (RHS, G, HPL) combination
// (both solvers, with and without the reverse sweep) and register the four kernels.
#pragma once
#include "ude_kernel.cuh"
#include "ude_registry.h"

namespace ude {
template <class R, int G, int HPL, int MINB, int SOLVER, int LK>
void register_one(const char* name, int priority) {
  kernel_registry().push_back(KernelEntry{R::KIND, R::D, R::NX, R::DIN, R::DOUT, G, HPL, SOLVER, LK, 1, Geo<R, G, HPL>::ROWS, priority,
                                          (const void*)ude_seg_kernel<R, G, HPL, SOLVER, LK, true, MINB>, name});
  kernel_registry().push_back(KernelEntry{R::KIND, R::D, R::NX, R::DIN, R::DOUT, G, HPL, SOLVER, LK, 0, Geo<R, G, HPL>::ROWS, priority,
                                          (const void*)ude_seg_kernel<R, G, HPL, SOLVER, LK, false, MINB>, name});
}
template <class R, int G, int HPL, int MINB>
void register_combo(const char* name, int priority) {
  register_one<R, G, HPL, MINB, 0, LK_STATE>(name, priority);
  register_one<R, G, HPL, MINB, 0, LK_TRAJ>(name, priority);
  register_one<R, G, HPL, MINB, 0, LK_DM>(name, priority);
  register_one<R, G, HPL, MINB, 1, LK_STATE>(name, priority);
  register_one<R, G, HPL, MINB, 1, LK_TRAJ>(name, priority);
  register_one<R, G, HPL, MINB, 1, LK_DM>(name, priority);
}
}  // namespace ude
