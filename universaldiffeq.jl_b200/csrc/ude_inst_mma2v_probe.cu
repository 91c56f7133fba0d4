// Tuning probes of the round-2 tensor-core kernel on the C3 shape (RK4, trajectory losses only); selected with UDE_KERNEL.
// The abl_* entries switch parts of the kernel OFF to attribute time (their results are wrong by construction; they carry
// the lowest priority and are never selected unless UDE_KERNEL names them).
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_mma2_probe<4, 1, 4, 1, 3, MMA2_PARK_GW>("node_d4x1_mma2_h32_p1_park", 9);
  register_mma2_probe<4, 1, 4, 1, 3, MMA2_PARK_GW | MMA2_ROT>("node_d4x1_mma2_h32_p1_rot", 9);
  register_mma2_probe<4, 1, 4, 1, 3, MMA2_PARK_GW | MMA2_TAB16 | MMA2_ROT>("node_d4x1_mma2_h32_p1_rt16", 9);
  register_mma2_probe<4, 1, 4, 1, 4, MMA2_PARK_GW>("node_d4x1_mma2_h32_p1_b4", 9);
  register_mma2_probe<4, 1, 4, 1, 3, MMA2_PARK_GW | MMA2_TAB16, 8>("node_d4x1_mma2_h32_p1_tb8", 9);
  register_mma2_probe<4, 1, 4, 1, 3, MMA2_PARK_GW | MMA2_TAB16, 2>("node_d4x1_mma2_h32_p1_tb2", 9);
});
}  // namespace
