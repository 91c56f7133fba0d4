// FP64 tensor-core (DMMA) kernels for the C3 shape: NODE d = 4, one covariate, hidden <= 32.
// Measured alternatives (C3, ms per evaluation): 3 blocks/SM, tanh in batches of 4: 8.5 (kept); 4 blocks/SM (128 regs):
// 9.0; tanh batches of 8: 9.0 (2 blocks) / 9.1 (3 blocks).
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_mma_combo<4, 1, 4, 3>("node_d4x1_mma_h32", 1);
});
}  // namespace
