// Host-side registry of the compiled sm_100a specialisations of ude_seg_kernel.
// Each ude_inst_*.cu translation unit registers its kernels at load time; ude_create picks one by
// (rhs_kind, d, n_cov, nn_in, nn_out) and the smallest padded hidden width G*HPL >= nn_hidden.
#pragma once
#include <vector>

namespace ude {

struct KernelEntry {
  int rhs_kind, d, n_cov, nn_in, nn_out;
  int dtype;               // UDE_F64 / UDE_F32: arithmetic type of the kernel
  int G, HPL, solver, lk, grad;   // lk: 0 state-space, 1 trajectory (shooting kinds), 2 derivative matching
  int rows_per_stage;      // stash rows (256 B each) per RK stage
  int elem_bytes;          // sizeof the arithmetic type (stash rows are 32 elements)
  int ws_bytes;            // per-warp shared-memory weight records (0: weights live in registers or in block fragments)
  int block_bytes;         // per-block dynamic shared memory besides the partial row (MMA kernels: weight fragments)
  int warp_bytes;          // per-warp dynamic shared memory (weight records + TMA staging ring + scratch tiles)
  int single_model_only;   // MMA kernels keep one model's weights per block
  int spw;                 // segments per task (32/G, or 8 for the MMA kernels)
  int warps_per_task;      // warps that cooperate on one task (MMA kernels with a split hidden layer: 4)
  int priority;            // lower = preferred when several (G, HPL) fit
  const void* fn;
  const char* name;
  int red_alias;           // 1: the block-partial row aliases the per-warp region (no shared memory of its own)
};

std::vector<KernelEntry>& kernel_registry();

struct RegisterKernels {
  explicit RegisterKernels(void (*fn)()) { fn(); }
};

}  // namespace ude
