// tuning variants of the C3 shape (selected with UDE_KERNEL=<name substring>)
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_combo<RhsNode<4, 1>, 8, 4, 3, true>("node_d4x1_g8h4_ws_b3", 6);
  register_combo<RhsNode<4, 1>, 16, 2, 3, true>("node_d4x1_g16h2_ws_b3", 7);
});
}  // namespace
