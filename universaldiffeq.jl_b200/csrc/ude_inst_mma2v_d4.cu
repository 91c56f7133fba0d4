// Round-2 FP64 tensor-core kernels (ude_kernel_mma2.cuh) for the C3 shape: NODE d = 4, one covariate, hidden <= 32.
// nt1: 8 segments per warp, 3 blocks per SM (the default); nt2: 16 segments per warp, 2 blocks per SM (measured slower:
// profiles/sweep_r02.md).
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_mma2_combo<4, 1, 4, 1, 3, MMA2_PARK_GW | MMA2_TAB16 | MMA2_AFFINE | MMA2_TANH3>("node_d4x1_mma2_h32_nt1", 0);
  register_mma2_combo<4, 1, 4, 2, 2, MMA2_PARK_GW | MMA2_TAB16 | MMA2_TANH3>("node_d4x1_mma2_h32_nt2", 2);
});
}  // namespace
