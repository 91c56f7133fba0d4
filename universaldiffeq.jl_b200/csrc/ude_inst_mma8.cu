// FP64 tensor-core (DMMA) kernels for the wide NODE of BASELINE config C5: d = 8, hidden 128.  The hidden layer is split
// over the four warps of a block (32 units each); the warps exchange their partial W2 h and W1^T z-bar through shared memory.
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  // BX = 1: b1 stays out of the MMAs (d = 8 fills the input slots exactly; ude_kernel_mma.cuh MmaGeo::BX): 16.0 -> 14.5 ms on C5
  // BX | RA = 3, 3 blocks per SM: the block-partial row (17.5 KB at H = 128) aliases the stage rings, 86 -> 69 KB per block, 168 registers:
  // 14.5 -> 13.5 ms on C5 (3 instead of 2 warps per scheduler)
  register_mma_combo<8, 0, 4, 3, 4, 4, 3>("node_d8_mma_h128", 0);
  register_mma_combo<8, 0, 4, 3, 1, 4, 1>("node_d8_mma_h32", 0);
  register_mma_combo<8, 0, 4, 2, 4, 4, 0>("node_d8_mma_h128_biasmma", 9);       // A/B: the bias as a constant-1 input slot
  register_mma_combo<8, 0, 4, 2, 4, 4, 1>("node_d8_mma_h128_b2", 9);            // A/B: 2 blocks per SM (250 registers, partial row in its own shared memory)
});
}  // namespace
