// FP64 tensor-core (DMMA) kernels for the wide NODE of BASELINE config C5: d = 8, hidden 128.  The hidden layer is split
// over the four warps of a block (32 units each); the warps exchange their partial W2 h and W1^T z-bar through shared memory.
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_mma_combo<8, 0, 4, 2, 4>("node_d8_mma_h128", 0);
  register_mma_combo<8, 0, 4, 3, 1>("node_d8_mma_h32", 0);
});
}  // namespace
