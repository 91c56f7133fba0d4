// Two lanes per segment, five hidden units per lane: the constructors' default hidden_units = 10 without padding
// (BASELINE configs C2 and C4), 16 segments per warp.  Measured +48 % (C2) / +38 % (C4) over the 4 x 3 mapping.
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_combo<RhsNode<2, 0>, 2, 5, 2>("node_d2_g2h5", 0);
  register_combo<RhsNodeTime<3, 0, 2>, 2, 5, 2>("node_time_d3_g2h5", 0);
  register_combo<RhsNode<4, 1>, 2, 5, 2>("node_d4x1_g2h5", 0);
  register_combo<RhsNode<2, 1>, 2, 5, 2>("node_d2x1_g2h5", 0);
});
}  // namespace
