"""ctypes binding of libude_cuda.so (include/ude_cuda.h) and the ``Engine`` wrapper that feeds it a Problem.

This is the Python counterpart of the `ccall` stub INTEGRATION.md gives for Julia: the same entry points, the same
argument order, no torch types.  There is no CPU fallback anywhere in this module: if the shared library is missing
or no sm_100 device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from .problem import MAX_SCALARS, Problem

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libude_cuda.so")
ABI_VERSION = 1
MAX_D = 8

UDE_OK = 0
STATUS = {0: "UDE_OK", -1: "UDE_ERR_INVALID", -2: "UDE_ERR_UNSUPPORTED", -3: "UDE_ERR_CUDA", -4: "UDE_ERR_NO_DEVICE",
          -5: "UDE_ERR_STATE", -6: "UDE_ERR_NCCL", -7: "UDE_NONFINITE"}
UDE_NONFINITE = -7

EXPORTS = ["ude_abi_version", "ude_device_count", "ude_create", "ude_destroy", "ude_last_error", "ude_set_data",
           "ude_set_model_weights", "ude_set_covariates", "ude_set_targets", "ude_set_skip", "ude_set_option", "ude_last_io_bytes", "ude_n_theta",
           "ude_n_models", "ude_steps_per_eval", "ude_kernel_name", "ude_launches_per_eval", "ude_loss_grad", "ude_loss_grad_device",
           "ude_forward", "ude_comm_unique_id", "ude_comm_init", "ude_adam", "ude_adam_used_graph", "ude_profile_begin", "ude_profile_read",
           "ude_measure_fp64_peak", "ude_measure_fp32_peak", "ude_peer_export", "ude_peer_attach",
           "ude_ukf_loss_grad", "ude_ukf_filter", "ude_forecast_chain", "ude_reg_smooth"]


class UkfOpts(C.Structure):
    """ude_ukf_opts (include/ude_cuda.h)"""
    _fields_ = [("alpha", C.c_double), ("beta", C.c_double), ("kappa", C.c_double), ("reg", C.c_double), ("Peta", C.c_double * 64)]


class UdeDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("struct_bytes", C.c_int32), ("device", C.c_int32), ("dtype", C.c_int32),
        ("solver", C.c_int32), ("substeps", C.c_int32),
        ("rhs_kind", C.c_int32), ("d", C.c_int32), ("n_cov", C.c_int32), ("nn_in", C.c_int32), ("nn_hidden", C.c_int32),
        ("nn_out", C.c_int32), ("n_scalars", C.c_int32), ("has_series_param", C.c_int32),
        ("off_w1", C.c_int64), ("off_b1", C.c_int64), ("off_w2", C.c_int64), ("off_b2", C.c_int64),
        ("off_scalar", C.c_int64 * MAX_SCALARS), ("off_series_param", C.c_int64), ("n_pm", C.c_int64),
        ("time_scale", C.c_double), ("time_shift", C.c_double),
        ("loss_kind", C.c_int32), ("pred_length", C.c_int32), ("proc_inv_dt", C.c_int32), ("shoot_from_theta", C.c_int32),
        ("shoot_first_scored", C.c_int32), ("remove_ends", C.c_int32),
        ("preroll_eps", C.c_double),
        ("A", C.c_double * (MAX_D * MAX_D)), ("B", C.c_double * (MAX_D * MAX_D)),
        ("lanes_per_segment", C.c_int32), ("hidden_per_lane", C.c_int32), ("weights_in_smem", C.c_int32), ("reserved", C.c_int32 * 5),
    ]


class UdeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


_LIB = None


def load_library():
    """Loads libude_cuda.so; raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(the CUDA library is the only implementation; there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dp = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p
    lib.ude_abi_version.restype = C.c_int
    lib.ude_device_count.restype = C.c_int
    lib.ude_create.restype = C.c_int
    lib.ude_create.argtypes = [C.POINTER(UdeDesc), C.POINTER(vp)]
    lib.ude_destroy.restype = C.c_int
    lib.ude_destroy.argtypes = [vp]
    lib.ude_last_error.restype = C.c_char_p
    lib.ude_last_error.argtypes = [vp]
    lib.ude_set_data.restype = C.c_int
    lib.ude_set_data.argtypes = [vp, i64, dp, dp, dp, i64, dp, dp]
    lib.ude_set_model_weights.restype = C.c_int
    lib.ude_set_model_weights.argtypes = [vp, dp, dp, dp, dp]
    lib.ude_set_covariates.restype = C.c_int
    lib.ude_set_covariates.argtypes = [vp, dp, dp, dp]
    lib.ude_set_targets.restype = C.c_int
    lib.ude_set_targets.argtypes = [vp, dp, dp]
    lib.ude_set_skip.restype = C.c_int
    lib.ude_set_skip.argtypes = [vp, dp]
    lib.ude_n_theta.restype = i64
    lib.ude_n_theta.argtypes = [vp]
    lib.ude_n_models.restype = i64
    lib.ude_n_models.argtypes = [vp]
    lib.ude_steps_per_eval.restype = i64
    lib.ude_steps_per_eval.argtypes = [vp]
    lib.ude_kernel_name.restype = C.c_char_p
    lib.ude_kernel_name.argtypes = [vp]
    lib.ude_launches_per_eval.restype = i32
    lib.ude_launches_per_eval.argtypes = [vp]
    lib.ude_loss_grad.restype = C.c_int
    lib.ude_loss_grad.argtypes = [vp, dp, i64, dp, dp]
    lib.ude_loss_grad_device.restype = C.c_int
    lib.ude_loss_grad_device.argtypes = [vp, vp, vp, vp, vp]
    lib.ude_forward.restype = C.c_int
    lib.ude_forward.argtypes = [vp, dp, i64, dp]
    lib.ude_comm_unique_id.restype = C.c_int
    lib.ude_comm_unique_id.argtypes = [vp, i32]
    lib.ude_comm_init.restype = C.c_int
    lib.ude_comm_init.argtypes = [vp, vp, i32, i32, i32]
    lib.ude_adam.restype = C.c_int
    lib.ude_adam.argtypes = [vp, dp, i64, C.c_double, i32, dp]
    lib.ude_adam_used_graph.restype = C.c_int
    lib.ude_adam_used_graph.argtypes = [vp]
    lib.ude_profile_begin.restype = C.c_int
    lib.ude_profile_begin.argtypes = [vp, i32]
    lib.ude_profile_read.restype = C.c_int
    lib.ude_profile_read.argtypes = [vp, dp, i32]
    lib.ude_measure_fp64_peak.restype = C.c_int
    lib.ude_measure_fp64_peak.argtypes = [i32, C.POINTER(C.c_double)]
    lib.ude_forecast_chain.restype = C.c_int
    lib.ude_forecast_chain.argtypes = [vp, dp, i64, dp, dp, C.c_double, C.c_double, dp]
    lib.ude_ukf_loss_grad.restype = C.c_int
    lib.ude_ukf_loss_grad.argtypes = [vp, C.POINTER(UkfOpts), dp, i64, dp, dp, dp, dp]
    lib.ude_ukf_filter.restype = C.c_int
    lib.ude_ukf_filter.argtypes = [vp, C.POINTER(UkfOpts), dp, i64, dp, dp, dp]
    lib.ude_peer_export.restype = C.c_int
    lib.ude_peer_export.argtypes = [vp, i32, vp, i32]
    lib.ude_peer_attach.restype = C.c_int
    lib.ude_peer_attach.argtypes = [vp, vp, i32, i32]
    lib.ude_measure_fp32_peak.restype = C.c_int
    lib.ude_measure_fp32_peak.argtypes = [i32, C.POINTER(C.c_double)]
    lib.ude_reg_smooth.restype = C.c_int
    lib.ude_reg_smooth.argtypes = [i32, i32, i32, i64, vp, vp, vp, vp, C.c_double, C.c_double, vp, vp, vp]
    lib.ude_set_option.restype = C.c_int
    lib.ude_set_option.argtypes = [vp, i32, i64]
    lib.ude_last_io_bytes.restype = C.c_int
    lib.ude_last_io_bytes.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    if lib.ude_abi_version() != ABI_VERSION:
        raise RuntimeError("libude_cuda ABI version mismatch")
    _LIB = lib
    return lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_desc(p: Problem, device: int = 0, lanes_per_segment: int = 0, hidden_per_lane: int = 0, weights_in_smem: int = 0,
              dtype: int = 0) -> UdeDesc:
    d = UdeDesc()
    d.abi_version = ABI_VERSION
    d.struct_bytes = C.sizeof(UdeDesc)
    d.device = device
    d.dtype = dtype            # 0 = UDE_F64, 1 = UDE_F32 (arithmetic in FP32, reductions / times / loss terms in FP64)
    for name in ("solver", "substeps", "rhs_kind", "d", "n_cov", "nn_in", "nn_hidden", "nn_out", "n_scalars",
                 "off_w1", "off_b1", "off_w2", "off_b2", "n_pm", "time_scale", "time_shift", "loss_kind", "pred_length",
                 "shoot_first_scored", "remove_ends", "preroll_eps"):
        setattr(d, name, getattr(p, name))
    d.has_series_param = int(p.has_series_param)
    d.proc_inv_dt = int(p.proc_inv_dt)
    d.shoot_from_theta = int(p.shoot_from_theta)
    d.off_series_param = int(p.off_series_param)
    for k, o in enumerate(p.off_scalar):
        d.off_scalar[k] = int(o)
    A = np.asfortranarray(p.A).reshape(-1, order="F")
    B = np.asfortranarray(p.B).reshape(-1, order="F")
    for k in range(p.d * p.d):
        d.A[k] = A[k]
        d.B[k] = B[k]
    d.lanes_per_segment = lanes_per_segment
    d.hidden_per_lane = hidden_per_lane
    d.weights_in_smem = weights_in_smem
    return d


def device_count() -> int:
    return int(load_library().ude_device_count())


def get_unique_id() -> bytes:
    lib = load_library()
    buf = C.create_string_buffer(128)
    rc = lib.ude_comm_unique_id(buf, 128)
    if rc != UDE_OK:
        raise UdeError(rc, lib.ude_last_error(None).decode())
    return buf.raw


def measure_fp64_peak(device: int = 0, f32: bool = False) -> float:
    lib = load_library()
    out = C.c_double(0.0)
    rc = (lib.ude_measure_fp32_peak if f32 else lib.ude_measure_fp64_peak)(device, C.byref(out))
    if rc != UDE_OK:
        raise UdeError(rc, lib.ude_last_error(None).decode())
    return float(out.value)


OPT_SPARSE_UHAT_IO = 1
OPT_UKF_CAUSAL_INTERVAL = 2


class Engine:
    """One ude_handle: the batched loss-and-gradient evaluator of a finalized Problem on one GPU."""

    def __init__(self, p: Problem, device: int = 0, lanes_per_segment: int = 0, hidden_per_lane: int = 0, weights_in_smem: int = 0,
                 dtype: int = 0):
        self.lib = load_library()
        self.problem = p
        self.handle = C.c_void_p()
        if p.rhs_kind >= 1000:          # a right-hand side written as expressions: its kernels live in a generated plugin
            from . import user_rhs
            rhs = p.user_rhs if getattr(p, "user_rhs", None) is not None else user_rhs.lookup(p.rhs_kind)
            if rhs is None:
                raise UdeError(-2, f"rhs_kind {p.rhs_kind} is not a catalog entry and no expression_rhs with that id exists in this process")
            user_rhs.ensure_kernels(rhs, p.nn_hidden, p.solver, dtype=dtype)
        desc = make_desc(p, device, lanes_per_segment, hidden_per_lane, weights_in_smem, dtype)
        rc = self.lib.ude_create(C.byref(desc), C.byref(self.handle))
        if rc != UDE_OK:
            raise UdeError(rc, self.lib.ude_last_error(None).decode())
        try:
            mos = p.model_of_series if p.n_models > 1 else None
            self._check(self.lib.ude_set_data(self.handle, p.n_series, _ptr(p.starts), _ptr(p.lengths), _ptr(mos),
                                              p.n_cols, _ptr(p.times), _ptr(p.data)))
            self._check(self.lib.ude_set_model_weights(self.handle, _ptr(p.obs_scale), _ptr(p.proc_scale),
                                                       _ptr(p.shoot_weight), _ptr(p.reg_weight)))
            if p.n_cov > 0:
                self._check(self.lib.ude_set_covariates(self.handle, _ptr(p.knot_off), _ptr(p.knot_t), _ptr(p.knot_x)))
            if p.dm_u is not None:
                self._check(self.lib.ude_set_targets(self.handle, _ptr(p.dm_u), _ptr(p.dm_dudt)))
            if p.skip_mask is not None:
                self._check(self.lib.ude_set_skip(self.handle, _ptr(p.skip_mask)))
            n = int(self.lib.ude_n_theta(self.handle))
            if n != p.n_theta:
                raise RuntimeError(f"theta layout mismatch: library {n}, host {p.n_theta}")
            # a catalog right-hand side at a shape (hidden width, many models) without a pre-compiled kernel: compile the same
            # hand-written trait for that geometry now and register it (user_rhs.ensure_catalog_kernels), then plan again
            if not self.lib.ude_kernel_name(self.handle) and p.rhs_kind < 1000 and lanes_per_segment == 0 and hidden_per_lane == 0 \
                    and weights_in_smem == 0 and os.environ.get("UDE_JIT", "1") != "0" and not os.environ.get("UDE_KERNEL"):
                msg = self.lib.ude_last_error(self.handle).decode()
                if "no compiled sm_100a kernel" in msg:
                    from . import user_rhs
                    try:
                        user_rhs.ensure_catalog_kernels(p, dtype)
                    except ValueError as ex:
                        raise UdeError(-2, f"{msg} [{ex}]")
                    if not self.lib.ude_kernel_name(self.handle):
                        raise UdeError(-2, self.lib.ude_last_error(self.handle).decode())
        except Exception:
            self.close()
            raise

    # -- plumbing --------------------------------------------------------------------------------
    def _check(self, rc, allow_nonfinite=False):
        if rc == UDE_OK or (allow_nonfinite and rc == UDE_NONFINITE):
            return rc
        raise UdeError(rc, self.lib.ude_last_error(self.handle).decode())

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.ude_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- the hot path ----------------------------------------------------------------------------
    @property
    def n_theta(self) -> int:
        return int(self.lib.ude_n_theta(self.handle))

    @property
    def n_models(self) -> int:
        return int(self.lib.ude_n_models(self.handle))

    def steps_per_eval(self) -> int:
        n = int(self.lib.ude_steps_per_eval(self.handle))
        if n < 0:
            raise UdeError(-1, self.lib.ude_last_error(self.handle).decode())
        return n

    def kernel_name(self) -> str:
        return self.lib.ude_kernel_name(self.handle).decode()

    def launches_per_eval(self) -> int:
        return int(self.lib.ude_launches_per_eval(self.handle))

    def loss_grad(self, theta: np.ndarray, want_grad: bool = True, loss_out=None, grad_out=None):
        """(loss[n_models], grad[n_theta] | None) through host buffers: H2D(theta) -> kernels -> D2H(loss, grad)."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        loss = np.empty(self.problem.n_models) if loss_out is None else loss_out
        grad = (np.empty(theta.shape[0]) if grad_out is None else grad_out) if want_grad else None
        self._check(self.lib.ude_loss_grad(self.handle, _ptr(theta), theta.shape[0], _ptr(loss), _ptr(grad)),
                    allow_nonfinite=True)
        return loss, grad

    def loss_grad_ptr(self, theta_ptr: int, n_theta: int, loss_ptr: int, grad_ptr: int):
        """Same call on raw HOST addresses (e.g. pinned torch tensors' data_ptr())."""
        return self._check(self.lib.ude_loss_grad(self.handle, theta_ptr, n_theta, loss_ptr, grad_ptr), allow_nonfinite=True)

    def loss_grad_device(self, d_theta: int, d_loss: int, d_grad: int, stream: int = 0):
        """Device-resident variant (raw device addresses); asynchronous on `stream` (0 = the handle's stream)."""
        return self._check(self.lib.ude_loss_grad_device(self.handle, d_theta, d_loss, d_grad or None, stream or None))

    def forward(self, theta: np.ndarray) -> np.ndarray:
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        out = np.zeros((self.problem.d, self.problem.n_cols), order="F")
        self._check(self.lib.ude_forward(self.handle, _ptr(theta), theta.shape[0], _ptr(out)))
        return out

    def adam(self, theta: np.ndarray, step_size: float, n_iter: int):
        """Device-resident Adam; returns (theta_new, loss_trace[n_iter, n_models])."""
        theta = np.array(theta, dtype=np.float64, order="C", copy=True)
        trace = np.zeros((max(n_iter, 1), self.problem.n_models))
        self._check(self.lib.ude_adam(self.handle, _ptr(theta), theta.shape[0], float(step_size), int(n_iter), _ptr(trace)))
        return theta, trace[:n_iter]

    def set_option(self, option: int, value: int = 1):
        """ude_set_option; OPT_SPARSE_UHAT_IO = 1: zero-copy block-start columns for (multiple-)shooting losses with
        page-locked host buffers -- the caller zeroes the gradient buffer once (include/ude_cuda.h)."""
        self._check(self.lib.ude_set_option(self.handle, int(option), int(value)))

    def last_io_bytes(self):
        """(host->device, device->host) bytes the last loss_grad call moved."""
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.ude_last_io_bytes(self.handle, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def profile_begin(self, max_evals: int):
        self._check(self.lib.ude_profile_begin(self.handle, int(max_evals)))

    def profile_read(self, capacity: int) -> np.ndarray:
        out = np.zeros(max(capacity, 1), dtype=np.float32)
        n = self.lib.ude_profile_read(self.handle, _ptr(out), int(capacity))
        if n < 0:
            self._check(n)
        return out[:n].astype(np.float64)

    def adam_used_graph(self) -> bool:
        return bool(self.lib.ude_adam_used_graph(self.handle))

    def forecast_chain(self, theta, u0, umax, umin, umean, l, rho):
        """ude_forecast_chain: the handle's one-interval series chained on the device; returns d x n_series."""
        th = np.ascontiguousarray(theta, dtype=np.float64)
        x0 = np.ascontiguousarray(u0, dtype=np.float64)
        lim = np.ascontiguousarray(np.concatenate([umax, umin, umean]), dtype=np.float64)
        out = np.zeros((self.problem.d, self.problem.n_series), order="F")
        self._check(self.lib.ude_forecast_chain(self.handle, th.ctypes.data, th.shape[0], x0.ctypes.data, lim.ctypes.data,
                                                float(l), float(rho), out.ctypes.data))
        return out

    def _ukf_opts(self, Peta, alpha, beta, kappa, reg):
        o = UkfOpts()
        o.alpha, o.beta, o.kappa, o.reg = float(alpha), float(beta), float(kappa), float(reg)
        d = self.problem.d
        Pe = np.asarray(Peta, dtype=np.float64).reshape(d, d)
        for j in range(d):
            for i in range(d):
                o.Peta[i + d * j] = Pe[i, j]
        return o

    def _ukf_call(self, fn):
        """The filter's auxiliary handle is a many-model one (a model per filter step): a shape whose only pre-compiled kernel
        is a single-model tensor-core kernel needs the lane-per-hidden-unit kernels compiled on demand (user_rhs)."""
        rc = fn()
        if rc == -2 and self.problem.rhs_kind < 1000 and os.environ.get("UDE_JIT", "1") != "0" \
                and "auxiliary handle: no compiled sm_100a kernel" in self.lib.ude_last_error(self.handle).decode():
            from . import user_rhs
            try:
                user_rhs.ensure_catalog_kernels(self.problem, 0)
            except ValueError:
                return rc
            rc = fn()
        return rc

    def _ukf_factor(self, pnu_factor):
        """(M, d, d) process-noise factors as column-major bytes per model; a single (d, d) factor is shared by every model"""
        d, M = self.problem.d, self.problem.n_models
        F = np.asarray(pnu_factor, dtype=np.float64)
        F = np.broadcast_to(F.reshape(-1, d, d), (M, d, d)) if F.size == d * d else F.reshape(M, d, d)
        return np.ascontiguousarray(F.transpose(0, 2, 1))

    def ukf_loss_grad(self, theta, pnu_factor, Peta, alpha=1e-3, beta=2.0, kappa=0.0, reg=0.0, want_grad=True, causal=False):
        """Unscented-Kalman-filter marginal likelihood (src/UnscentedKalmanFilter.jl:76-87): returns
        (loss, grad_theta, grad_pnu_factor) with P_nu = F F', F = pnu_factor (d x d).  causal=False (default) integrates the
        sigma points over [times[t], times[t] + dt] as the reference does (:83, :93); causal=True over [times[t-1], times[t]].
        Many-model handles: pnu_factor (n_models, d, d) (or one (d, d) for all), loss (n_models,), grad_pnu_factor
        (n_models, d, d); single-model handles keep the scalar / (d, d) shapes."""
        self.set_option(OPT_UKF_CAUSAL_INTERVAL, int(bool(causal)))
        d, M = self.problem.d, self.problem.n_models
        th = np.ascontiguousarray(theta, dtype=np.float64)
        F = self._ukf_factor(pnu_factor)
        o = self._ukf_opts(Peta, alpha, beta, kappa, reg)
        loss = np.zeros(M)
        g = np.zeros_like(th) if want_grad else None
        gF = np.zeros((M, d, d)) if want_grad else None
        rc = self._ukf_call(lambda: self.lib.ude_ukf_loss_grad(self.handle, C.byref(o), th.ctypes.data, th.shape[0], F.ctypes.data, loss.ctypes.data,
                                                               g.ctypes.data if want_grad else None, gF.ctypes.data if want_grad else None))
        if rc != UDE_OK and rc != UDE_NONFINITE:
            self._check(rc)
        if want_grad:
            gF = gF.transpose(0, 2, 1).copy()
        if M == 1:
            return float(loss[0]), g, (gF[0] if want_grad else None)
        return loss, g, gF

    def ukf_filter(self, theta, pnu_factor, Peta, alpha=1e-3, beta=2.0, kappa=0.0, causal=False):
        """ukf_smoothing (src/UnscentedKalmanFilter.jl:59-72): filtered states (d x n_cols) and covariances (d x d x n_cols)."""
        self.set_option(OPT_UKF_CAUSAL_INTERVAL, int(bool(causal)))
        d = self.problem.d
        th = np.ascontiguousarray(theta, dtype=np.float64)
        F = self._ukf_factor(pnu_factor)
        o = self._ukf_opts(Peta, alpha, beta, kappa, 0.0)
        xs = np.zeros((d, self.problem.n_cols), order="F")
        Ps = np.zeros((d, d, self.problem.n_cols), order="F")
        rc = self._ukf_call(lambda: self.lib.ude_ukf_filter(self.handle, C.byref(o), th.ctypes.data, th.shape[0], F.ctypes.data, xs.ctypes.data,
                                                            Ps.ctypes.data))
        if rc != UDE_OK and rc != UDE_NONFINITE:
            self._check(rc)
        return xs, Ps

    def peer_export(self, n_ranks: int) -> bytes:
        """Allocate this rank's exchange buffer; returns the 128-byte blob the other ranks attach with."""
        buf = C.create_string_buffer(128)
        self._check(self.lib.ude_peer_export(self.handle, n_ranks, buf, 128))
        return buf.raw

    def peer_attach(self, blobs, n_ranks: int, rank: int):
        """blobs: the n_ranks exported blobs in rank order.  From here on loss_grad / adam finish with the fused
        finalize + peer-memory exchange kernel instead of an NCCL allreduce."""
        raw = b"".join(bytes(b) for b in blobs)
        if len(raw) != 128 * n_ranks:
            raise ValueError("need n_ranks blobs of 128 bytes")
        buf = C.create_string_buffer(raw, len(raw))
        self._check(self.lib.ude_peer_attach(self.handle, buf, n_ranks, rank))

    def comm_init(self, unique_id: bytes, n_ranks: int, rank: int):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self.lib.ude_comm_init(self.handle, buf, 128, n_ranks, rank))
