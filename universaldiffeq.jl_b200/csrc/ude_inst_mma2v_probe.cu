// Tuning probes of the round-2 tensor-core kernel on the C3 shape (RK4, trajectory losses only); selected with UDE_KERNEL.
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_mma2_probe<4, 1, 4, 1, 3, MMA2_PARK_GW | MMA2_TAB16>("node_d4x1_mma2_h32_p1_noaff", 9);
  register_mma2_probe<4, 1, 4, 1, 3, MMA2_PARK_GW | MMA2_TAB16 | MMA2_SMEM_PF>("node_d4x1_mma2_h32_p1_spf", 9);
  register_mma2_probe<4, 1, 4, 1, 3, MMA2_PARK_GW | MMA2_TAB16 | MMA2_AFFINE>("node_d4x1_mma2_h32_p1_tanh1", 9);     // round-2a default: 14-instruction tanh
  register_mma2_probe<4, 1, 4, 1, 3, MMA2_PARK_GW | MMA2_TAB16 | MMA2_AFFINE | MMA2_TANH3 | MMA2_RING3>("node_d4x1_mma2_h32_p1_ring3", 9);
});
}  // namespace
