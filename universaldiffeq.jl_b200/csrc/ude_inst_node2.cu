// NODE d = 2 (BASELINE config C2: NN 2 -> 10 -> 2) and d = 2 with one covariate.
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_combo<RhsNode<2, 0>, 4, 3, 4>("node_d2_g4h3", 0);
  register_combo<RhsNode<2, 0>, 16, 2, 4>("node_d2_g16h2", 0);
  register_combo<RhsNode<2, 1>, 4, 3, 4>("node_d2x1_g4h3", 0);
});
}  // namespace
