// Helper for the ude_inst_*.cu translation units: instantiate ude_seg_kernel for one (RHS, G, HPL) combination
// (both solvers, with and without the reverse sweep) and register the four kernels.
#pragma once
#include "ude_kernel.cuh"
#include "ude_registry.h"

namespace ude {
template <class R, int G, int HPL, int MINB>
void register_combo(const char* name, int priority) {
  auto add = [&](int solver, int grad, const void* fn) {
    kernel_registry().push_back(KernelEntry{R::KIND, R::D, R::NX, R::DIN, R::DOUT, G, HPL, solver, grad,
                                            Geo<R, G, HPL>::ROWS, priority, fn, name});
  };
  add(0, 1, (const void*)ude_seg_kernel<R, G, HPL, 0, true, MINB>);
  add(0, 0, (const void*)ude_seg_kernel<R, G, HPL, 0, false, MINB>);
  add(1, 1, (const void*)ude_seg_kernel<R, G, HPL, 1, true, MINB>);
  add(1, 0, (const void*)ude_seg_kernel<R, G, HPL, 1, false, MINB>);
}
}  // namespace ude
