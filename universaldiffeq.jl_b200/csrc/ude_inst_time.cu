// Time-dependent NODEs built with CustomDerivatives (docs/src/examples.md:33-53; BASELINE config C4: coral UDE).
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_combo<RhsNodeTime<3, 0, 2>, 4, 3, 4>("node_time_d3_g4h3", 0);
  register_combo<RhsNodeTime<2, 1, 2>, 4, 3, 4>("node_time_d2x1_g4h3", 0);
});
}  // namespace
