// Helper for the ude_inst_*.cu translation units: instantiate the kernels of one (RHS, G, HPL, weight storage)
// combination -- both solvers, the three loss families, with and without the reverse sweep -- and register them.
#pragma once
#include "ude_kernel.cuh"
#include "ude_kernel_mma.cuh"
#include "ude_kernel_mma2.cuh"
#include "ude_kernel_tf32.cuh"
#include "ude_registry.h"

namespace ude {
// the two typed copies of the lane-per-hidden-unit family
struct F64 {
  using real = double;
  static constexpr int DTYPE = UDE_F64;
  template <class R, int G, int HPL> using Geo = f64::Geo<R, G, HPL>;
  template <class R, int G, int HPL, int SOLVER, int LK, bool GRAD, int MINB, bool WS>
  static const void* kernel() { return (const void*)f64::ude_seg_kernel<R, G, HPL, SOLVER, LK, GRAD, MINB, WS>; }
};
struct F32 {
  using real = float;
  static constexpr int DTYPE = UDE_F32;
  template <class R, int G, int HPL> using Geo = f32::Geo<R, G, HPL>;
  template <class R, int G, int HPL, int SOLVER, int LK, bool GRAD, int MINB, bool WS>
  static const void* kernel() { return (const void*)f32::ude_seg_kernel<R, G, HPL, SOLVER, LK, GRAD, MINB, WS>; }
};

template <class NS, class R, int G, int HPL, int MINB, int SOLVER, int LK, bool WS>
void register_one(const char* name, int priority) {
  using GG = typename NS::template Geo<R, G, HPL>;
  const int eb = (int)sizeof(typename NS::real);
  const int wsb = WS ? G * GG::WREC2 * 2 * eb : 0;
  const int ring = (LK != LK_DM) ? 2 * Tab<SOLVER>::S * GG::ROWS * 32 * eb : 0;
  kernel_registry().push_back(KernelEntry{R::KIND, R::D, R::NX, R::DIN, R::DOUT, NS::DTYPE, G, HPL, SOLVER, LK, 1, GG::ROWS, eb, wsb, 0,
                                          wsb + ring, 0, 32 / G, 1, priority,
                                          NS::template kernel<R, G, HPL, SOLVER, LK, true, MINB, WS>(), name});
  kernel_registry().push_back(KernelEntry{R::KIND, R::D, R::NX, R::DIN, R::DOUT, NS::DTYPE, G, HPL, SOLVER, LK, 0, GG::ROWS, eb, wsb, 0,
                                          wsb, 0, 32 / G, 1, priority,
                                          NS::template kernel<R, G, HPL, SOLVER, LK, false, MINB, WS>(), name});
}
template <class NS, class R, int G, int HPL, int MINB, bool WS = false>
void register_combo_t(const char* name, int priority) {
  register_one<NS, R, G, HPL, MINB, 0, LK_STATE, WS>(name, priority);
  register_one<NS, R, G, HPL, MINB, 0, LK_TRAJ, WS>(name, priority);
  register_one<NS, R, G, HPL, MINB, 0, LK_DM, WS>(name, priority);
  register_one<NS, R, G, HPL, MINB, 1, LK_STATE, WS>(name, priority);
  register_one<NS, R, G, HPL, MINB, 1, LK_TRAJ, WS>(name, priority);
  register_one<NS, R, G, HPL, MINB, 1, LK_DM, WS>(name, priority);
  if constexpr (NS::DTYPE == UDE_F64 && !WS) register_one<NS, R, G, HPL, MINB, 2, LK_STATE, WS>(name, priority);   // Euler
}
// FP64 (R is an f64:: trait, found through `using namespace f64`)
template <class R, int G, int HPL, int MINB, bool WS = false>
void register_combo(const char* name, int priority) { register_combo_t<F64, R, G, HPL, MINB, WS>(name, priority); }
// FP32 (pass the f32:: trait, e.g. register_combo_f32<f32::RhsNode<4, 1>, 8, 4, 3>(...))
template <class R, int G, int HPL, int MINB, bool WS = false>
void register_combo_f32(const char* name, int priority) { register_combo_t<F32, R, G, HPL, MINB, WS>(name, priority); }

// tensor-core (FP64 DMMA) family: NODE right-hand sides, 8 segments per task, HW warps per task each owning HT hidden
// tiles (hidden width padded to 8*HT*HW)
template <int D, int NX, int HT, int HW, int MINB, int SOLVER, int LK, int TB = 4, int BX = 0>
void register_mma_one(const char* name, int priority) {
  using MG = MmaGeo<D, NX, HT, HW, TB, BX>;
  // registry fields G / HPL are reused as "segments per warp-task = 32/G" and capacity G*HPL >= nn_hidden
  const int G = 4, HPLcap = 2 * HT * HW;
  kernel_registry().push_back(KernelEntry{UDE_RHS_NODE, D, NX, D + NX, D, UDE_F64, G, HPLcap, SOLVER, LK, 1, MG::REC / 32, 8, 0,
                                          MmaSmem<MG, LK, true>::BLOCK * 8, MmaSmem<MG, LK, true>::WARP * 8, 1, 8, HW, priority,
                                          (const void*)ude_mma_kernel<D, NX, HT, HW, SOLVER, LK, true, MINB, TB, BX>, name, MG::RA ? 1 : 0});
  kernel_registry().push_back(KernelEntry{UDE_RHS_NODE, D, NX, D + NX, D, UDE_F64, G, HPLcap, SOLVER, LK, 0, MG::REC / 32, 8, 0,
                                          MmaSmem<MG, LK, false>::BLOCK * 8, MmaSmem<MG, LK, false>::WARP * 8, 1, 8, HW, priority,
                                          (const void*)ude_mma_kernel<D, NX, HT, HW, SOLVER, LK, false, MINB, TB, BX>, name});
}
template <int D, int NX, int HT, int MINB, int HW = 1, int TB = 4, int BX = 0>
void register_mma_combo(const char* name, int priority) {
  register_mma_one<D, NX, HT, HW, MINB, 0, LK_STATE, TB, BX>(name, priority);
  register_mma_one<D, NX, HT, HW, MINB, 0, LK_TRAJ, TB, BX>(name, priority);
  register_mma_one<D, NX, HT, HW, MINB, 0, LK_DM, TB, BX>(name, priority);
  register_mma_one<D, NX, HT, HW, MINB, 1, LK_STATE, TB, BX>(name, priority);
  register_mma_one<D, NX, HT, HW, MINB, 1, LK_TRAJ, TB, BX>(name, priority);
  register_mma_one<D, NX, HT, HW, MINB, 1, LK_DM, TB, BX>(name, priority);
}

// round-2 tensor-core family (ude_kernel_mma2.cuh): NT M-tiles (8 segments each) per warp, state-space and trajectory
// loss families (derivative matching keeps ude_mma_kernel); OPT = MMA2_* bits, TB = tanh chains interleaved per batch
template <int D, int NX, int HT, int NT, int MINB, int SOLVER, int LK, int OPT, int TB>
void register_mma2_one(const char* name, int priority) {
  using MG = Mma2Geo<D, NX, HT, NT, OPT, TB>;
  const int G = 4, HPLcap = 2 * HT;
  kernel_registry().push_back(KernelEntry{UDE_RHS_NODE, D, NX, D + NX, D, UDE_F64, G, HPLcap, SOLVER, LK, 1, MG::RECW / 32, 8, 0,
                                          MG::NFRAG * 32 * 8, MG::WARP_GRAD * 8, 1, 8 * NT, 1, priority,
                                          (const void*)ude_mma2_kernel<D, NX, HT, NT, SOLVER, LK, true, MINB, OPT, TB>, name});
  kernel_registry().push_back(KernelEntry{UDE_RHS_NODE, D, NX, D + NX, D, UDE_F64, G, HPLcap, SOLVER, LK, 0, MG::RECW / 32, 8, 0,
                                          MG::NFRAG * 32 * 8, MG::WARP_FWD * 8, 1, 8 * NT, 1, priority,
                                          (const void*)ude_mma2_kernel<D, NX, HT, NT, SOLVER, LK, false, MINB, OPT, TB>, name});
}
template <int D, int NX, int HT, int NT, int MINB, int OPT = 0, int TB = 4>
void register_mma2_combo(const char* name, int priority) {
  register_mma2_one<D, NX, HT, NT, MINB, 0, LK_STATE, OPT, TB>(name, priority);
  register_mma2_one<D, NX, HT, NT, MINB, 0, LK_TRAJ, OPT, TB>(name, priority);
  register_mma2_one<D, NX, HT, NT, MINB, 1, LK_STATE, OPT, TB>(name, priority);
  register_mma2_one<D, NX, HT, NT, MINB, 1, LK_TRAJ, OPT, TB>(name, priority);
}
// tuning variants: RK4 + trajectory family only (what the C3 sweep measures)
template <int D, int NX, int HT, int NT, int MINB, int OPT = 0, int TB = 4>
void register_mma2_probe(const char* name, int priority) {
  register_mma2_one<D, NX, HT, NT, MINB, 0, LK_TRAJ, OPT, TB>(name, priority);
}

// FP32-mode tensor-core kernel (ude_kernel_tf32.cuh): NODE d = 8, hidden <= 32 * HTW, 16 segments per block-task, the hidden
// layer split over the block's four warps; state-space and trajectory loss families
template <int D, int HTW, int MINB, int SOLVER, int LK>
void register_tf32_one(const char* name, int priority) {
  using TG = Tf32Geo<D, HTW>;
  const int G = 4, HPLcap = 2 * HTW * TG::HW;
  kernel_registry().push_back(KernelEntry{UDE_RHS_NODE, D, 0, D, D, UDE_F32, G, HPLcap, SOLVER, LK, 1, TG::REC / 32, 4, 0,
                                          TG::BLOCK_BYTES, TG::WARP_F * 4, 1, 16, TG::HW, priority,
                                          (const void*)ude_tf32_kernel<D, HTW, SOLVER, LK, true, MINB>, name});
  kernel_registry().push_back(KernelEntry{UDE_RHS_NODE, D, 0, D, D, UDE_F32, G, HPLcap, SOLVER, LK, 0, TG::REC / 32, 4, 0,
                                          TG::BLOCK_BYTES, 0, 1, 16, TG::HW, priority,
                                          (const void*)ude_tf32_kernel<D, HTW, SOLVER, LK, false, MINB>, name});
}
template <int D, int HTW, int MINB>
void register_tf32_combo(const char* name, int priority) {
  register_tf32_one<D, HTW, MINB, 0, LK_STATE>(name, priority);
  register_tf32_one<D, HTW, MINB, 0, LK_TRAJ>(name, priority);
  register_tf32_one<D, HTW, MINB, 1, LK_STATE>(name, priority);
  register_tf32_one<D, HTW, MINB, 1, LK_TRAJ>(name, priority);
}
}  // namespace ude
