// Host-side registry of the compiled sm_100a specialisations of ude_seg_kernel.
// Each ude_inst_*.cu translation unit registers its kernels at load time; ude_create picks one by
// (rhs_kind, d, n_cov, nn_in, nn_out) and the smallest padded hidden width G*HPL >= nn_hidden.
#pragma once
#include <vector>

namespace ude {

struct KernelEntry {
  int rhs_kind, d, n_cov, nn_in, nn_out;
  int G, HPL, solver, lk, grad;   // lk: 0 state-space, 1 trajectory (shooting kinds), 2 derivative matching
  int rows_per_stage;      // stash rows (256 B each) per RK stage
  int ws_doubles;          // per-warp shared-memory weight records (0: weights live in registers)
  int stage_ring;          // 1: the kernel uses the per-warp two-step TMA staging ring
  int priority;            // lower = preferred when several (G, HPL) fit
  const void* fn;
  const char* name;
};

std::vector<KernelEntry>& kernel_registry();

struct RegisterKernels {
  explicit RegisterKernels(void (*fn)()) { fn(); }
};

}  // namespace ude
