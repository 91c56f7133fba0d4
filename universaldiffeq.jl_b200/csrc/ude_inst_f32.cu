// Optional FP32 mode (BASELINE north star: 1e-4 parity bar): the lane-per-hidden-unit family with float arithmetic.
// Times, covariates, loss terms, adjoint seeds and every reduction stay in FP64.
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_combo_f32<f32::RhsNode<2, 0>, 2, 5, 3>("f32_node_d2_g2h5", 0);
  register_combo_f32<f32::RhsNode<4, 1>, 2, 5, 3>("f32_node_d4x1_g2h5", 0);
  register_combo_f32<f32::RhsNode<4, 1>, 8, 4, 3>("f32_node_d4x1_g8h4", 1);
  register_combo_f32<f32::RhsNodeTime<3, 0, 2>, 2, 5, 3>("f32_node_time_d3_g2h5", 0);
});
}  // namespace
