"""Compile (here, without a GPU) the on-demand kernel plugins the GPU tests and bench.py ask for, so that they travel to the
GPU box with the tree instead of being compiled there (either works: nvcc is on the box too).
usage: python tools/prebuild_test_plugins.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import ude_b200 as ude
    import cases
    import user_rhs_cases as urc
    U = ude.user_rhs
    t0 = time.time()
    for rhs, h in [("lv_ude", 20), ("node_time_d3", 24), ("fishery", 32), ("lv_ude_abs", 16), ("series_growth_d1", 16), ("node_d4", 100),
                   ("node_d2x2", 32)]:                                       # tests/test_gpu_parity.py::test_catalog_shapes_compiled_on_demand
        prob, _ = cases.make_case(rhs, "cond_lik", 0, hidden=h, lengths=(5, 3))
        U.ensure_catalog_kernels(prob, 0)
    lv, ep, gr = urc.lv_twin(ude), urc.epidemic(ude), urc.growth_twin(ude)
    for r, h, s in [(lv, 10, 0), (lv, 10, 1), (lv, 6, 0), (ep, 6, 0), (ep, 6, 1), (ep, 12, 0), (ep, 20, 1), (gr, 4, 0), (gr, 4, 1)]:
        U.ensure_kernels(r, h, s)                                            # tests/test_gpu_user_rhs.py, bench.py user_rhs_lv
    U.ensure_kernels(lv, 10, 0, dtype=1); U.ensure_kernels(ep, 6, 0, dtype=1)     # test_user_rhs_fp32_mode
    for m, h in ((3, 6), (10, 10)):
        model, _, _ = urc.one_hot_model(ude, m=m, n_pts=6, hidden=h)
        U.ensure_kernels(model.rhs, h, 0)
    print(f"{len(os.listdir(U._CACHE)) // 2} plugins in {U._CACHE} ({time.time() - t0:.0f} s)")


if __name__ == "__main__":
    main()
