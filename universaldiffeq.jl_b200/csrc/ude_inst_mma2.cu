// FP64 tensor-core (DMMA) kernels for d = 2 NODEs (BASELINE config C2: NN 2 -> 10 -> 2, hidden padded to 16).
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_mma_combo<2, 0, 2, 4>("node_d2_mma_h16", 3);
});
}  // namespace
