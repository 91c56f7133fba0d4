// FP32 mode, BASELINE config C5 (wide NODE d = 8, hidden 128): 3xTF32 tensor-core kernel (ude_kernel_tf32.cuh).
#include "ude_inst.cuh"
namespace {
using namespace ude;
RegisterKernels reg([] {
  register_tf32_combo<8, 4, 2>("f32_node_d8_tf32_h128", -1);
});
}  // namespace
