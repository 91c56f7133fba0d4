// FP64 tensor-core (DMMA) kernels with the hidden layer split over the block's four warps: up to 128 hidden units for
// d = 2 and d = 4 NODEs (with and without one covariate).
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_mma_combo<4, 1, 4, 2, 4>("node_d4x1_mma_h128", 4);
  register_mma_combo<4, 0, 4, 2, 4, 4, 1>("node_d4_mma_h128", 4);
  register_mma_combo<2, 0, 4, 2, 4>("node_d2_mma_h128", 4);
});
}  // namespace
