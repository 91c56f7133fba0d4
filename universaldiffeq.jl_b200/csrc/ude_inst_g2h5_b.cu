// 2 x 5 mapping (hidden_units = 10) for the remaining NODE shapes and the parametric UDE right-hand sides.
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_combo<RhsNode<1, 0>, 2, 5, 2>("node_d1_g2h5", 0);
  register_combo<RhsNode<3, 0>, 2, 5, 2>("node_d3_g2h5", 0);
  register_combo<RhsNode<4, 0>, 2, 5, 2>("node_d4_g2h5", 0);
  register_combo<RhsLvUdeAbs, 2, 5, 2>("lv_ude_abs_g2h5", 0);
  register_combo<RhsLvUde<0>, 2, 5, 2>("lv_ude_g2h5", 0);
});
}  // namespace
