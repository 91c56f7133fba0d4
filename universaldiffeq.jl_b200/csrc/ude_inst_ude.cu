// Parametric UDE right-hand sides of the reference's README / tests / docs.
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_combo<RhsLvUdeAbs, 4, 3, 4>("lv_ude_abs_g4h3", 1);
  register_combo<RhsLvUde<0>, 4, 3, 4>("lv_ude_g4h3", 0);
  register_combo<RhsLvUde<1>, 4, 3, 4>("lv_ude_x1_g4h3", 0);
  register_combo<RhsFishery, 4, 3, 4>("fishery_g4h3", 1);
  register_combo<RhsNnDecay<1>, 4, 3, 4>("nn_decay_d1_g4h3", 0);
  register_combo<RhsSeriesGrowth<1>, 4, 3, 4>("series_growth_d1_g4h3", 0);
});
}  // namespace
