// Wide NODE d = 8 (BASELINE config C5: NN 8 -> 128 -> 8).
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_combo<RhsNode<8, 0>, 32, 4, 2>("node_d8_g32h4", 2);
  register_combo<RhsNode<8, 0>, 16, 2, 2>("node_d8_g16h2", 2);
});
}  // namespace
