// C3 shape with the lane's weights in shared memory (gradient accumulators stay in registers).
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_combo<RhsNode<4, 1>, 8, 4, 2, true>("node_d4x1_g8h4_ws_b2", 2);
  register_combo<RhsNode<4, 1>, 16, 2, 4, true>("node_d4x1_g16h2_ws_b4", 5);
});
}  // namespace
