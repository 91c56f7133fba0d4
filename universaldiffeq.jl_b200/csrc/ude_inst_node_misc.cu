// NODE / MultiNODE shapes beyond the BASELINE configs: d = 1, 3, 4 without covariates, d = 1, 3 with one covariate,
// d = 2 with two; hidden width <= 12 (the constructors' default hidden_units = 10) and <= 32.
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_combo<RhsNode<1, 0>, 4, 3, 4>("node_d1_g4h3", 0);
  register_combo<RhsNode<3, 0>, 4, 3, 4>("node_d3_g4h3", 0);
  register_combo<RhsNode<4, 0>, 4, 3, 4>("node_d4_g4h3", 0);
  register_combo<RhsNode<1, 1>, 4, 3, 4>("node_d1x1_g4h3", 0);
  register_combo<RhsNode<3, 1>, 4, 3, 4>("node_d3x1_g4h3", 0);
  register_combo<RhsNode<2, 2>, 4, 3, 4>("node_d2x2_g4h3", 0);
  register_combo<RhsNode<3, 0>, 16, 2, 3>("node_d3_g16h2", 1);
  register_mma_combo<4, 0, 4, 3, 1, 4, 1>("node_d4_mma_h32", 1);      // BX: bias outside the MMAs (MmaGeo::BX)
});
}  // namespace
