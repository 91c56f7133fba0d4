// FP32 mode, wide NODE of BASELINE config C5 (d = 8, hidden 128) and the README's Lotka-Volterra UDE.
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_combo_f32<f32::RhsNode<8, 0>, 32, 4, 3>("f32_node_d8_g32h4", 0);
  register_combo_f32<f32::RhsNode<8, 0>, 16, 2, 3>("f32_node_d8_g16h2", 1);
  register_combo_f32<f32::RhsLvUdeAbs, 2, 5, 3>("f32_lv_ude_abs_g2h5", 0);
  register_combo_f32<f32::RhsLvUde<0>, 2, 5, 3>("f32_lv_ude_g2h5", 0);
});
}  // namespace
