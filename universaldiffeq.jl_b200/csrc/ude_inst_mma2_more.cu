// Round-2 tensor-core kernel (ude_kernel_mma2.cuh) for the plain NODE shape d = 4 with up to 32 hidden units (no covariate): the
// bias is the only extra input slot, so MMA2_AFFINE drops the second layer-1 k-step entirely.  6.92 vs 7.50 ms for the round-1
// kernel on C3 without its covariate (profiles/node_shape_sweep_r02.json); d = 8, H = 32 measured equal to the round-1 kernel
// (one-interval segments), so it keeps that one.
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_mma2_combo<4, 0, 4, 1, 3, MMA2_PARK_GW | MMA2_TAB16 | MMA2_AFFINE | MMA2_TANH3>("node_d4_mma2_h32_nt1", 0);
});
}  // namespace
