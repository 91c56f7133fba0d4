// FP64 tensor-core (DMMA) kernels for the C3 shape: NODE d = 4, one covariate, hidden <= 32.
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_mma_combo<4, 1, 4, 3>("node_d4x1_mma_h32", 1);
  register_mma_combo<4, 1, 4, 4>("node_d4x1_mma_h32_b4", 8);
});
}  // namespace
