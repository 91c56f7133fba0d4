// MultiNODE with covariates, d = 4 (BASELINE config C3: NN 5 -> 32 -> 4) and plain d = 4 NODE.
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_combo<RhsNode<4, 1>, 16, 2, 3>("node_d4x1_g16h2", 4);
  register_combo<RhsNode<4, 1>, 8, 4, 2>("node_d4x1_g8h4", 3);
  register_combo<RhsNode<4, 1>, 4, 3, 3>("node_d4x1_g4h3", 0);
});
}  // namespace
