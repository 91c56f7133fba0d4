// FP64 tensor-core (DMMA) kernels for NODE / MultiNODE shapes beyond the BASELINE configs: d = 1..3 (with and without
// one covariate) up to 32 hidden units.
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_mma_combo<1, 0, 4, 3>("node_d1_mma_h32", 3);
  register_mma_combo<2, 0, 4, 3>("node_d2_mma_h32", 3);
  register_mma_combo<3, 0, 4, 3>("node_d3_mma_h32", 3);
  register_mma_combo<1, 1, 4, 3>("node_d1x1_mma_h32", 3);
  register_mma_combo<2, 1, 4, 3>("node_d2x1_mma_h32", 3);
  register_mma_combo<3, 1, 4, 3, 1, 4, 1>("node_d3x1_mma_h32", 3);     // BX: 3 + 1 inputs fill one k-step
});
}  // namespace
