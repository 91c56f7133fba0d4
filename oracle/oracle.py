"""ctypes wrapper around oracle/libude_oracle.so -- TEST INFRASTRUCTURE ONLY (see ude_oracle.h).

PARITY UNPINNED: the Julia reference cannot run here; see the header of ude_oracle.h.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _OrcProblem(C.Structure):
    _fields_ = [
        ("d", C.c_int32), ("n_cov", C.c_int32), ("nn_in", C.c_int32), ("nn_hidden", C.c_int32), ("nn_out", C.c_int32),
        ("rhs_kind", C.c_int32), ("n_scalars", C.c_int32), ("has_series_param", C.c_int32),
        ("solver", C.c_int32), ("substeps", C.c_int32),
        ("off_w1", C.c_int64), ("off_b1", C.c_int64), ("off_w2", C.c_int64), ("off_b2", C.c_int64),
        ("off_scalar", C.c_int64 * 8), ("off_series_param", C.c_int64), ("n_pm", C.c_int64),
        ("time_scale", C.c_double), ("time_shift", C.c_double),
        ("loss_kind", C.c_int32), ("pred_length", C.c_int32), ("proc_inv_dt", C.c_int32),
        ("shoot_from_theta", C.c_int32), ("shoot_first_scored", C.c_int32), ("remove_ends", C.c_int32),
        ("preroll_eps", C.c_double),
        ("A", C.c_void_p), ("B", C.c_void_p),
        ("n_series", C.c_int64), ("n_cols", C.c_int64), ("n_models", C.c_int64), ("n_theta", C.c_int64),
        ("starts", C.c_void_p), ("lengths", C.c_void_p), ("model_of_series", C.c_void_p),
        ("times", C.c_void_p), ("data", C.c_void_p),
        ("theta_base", C.c_void_p), ("model_col0", C.c_void_p), ("model_series0", C.c_void_p),
        ("obs_scale", C.c_void_p), ("proc_scale", C.c_void_p), ("shoot_weight", C.c_void_p), ("reg_weight", C.c_void_p),
        ("knot_off", C.c_void_p), ("knot_t", C.c_void_p), ("knot_x", C.c_void_p),
        ("dm_u", C.c_void_p), ("dm_dudt", C.c_void_p), ("skip_mask", C.c_void_p),
    ]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libude_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("ude_oracle.c", "ude_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libude_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_loss_grad.restype = C.c_int
        _LIB.orc_loss_grad.argtypes = [C.POINTER(_OrcProblem), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _LIB.orc_forward.restype = C.c_int
        _LIB.orc_forward.argtypes = [C.POINTER(_OrcProblem), C.c_void_p, C.c_void_p, C.c_int]
        _LIB.orc_count_steps.restype = C.c_int64
        _LIB.orc_count_steps.argtypes = [C.POINTER(_OrcProblem)]
        _LIB.orc_max_threads.restype = C.c_int
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _marshal(p):
    """p: universaldiffeq.jl_b200.problem.Problem (finalized).  Returns (struct, keepalive)."""
    s = _OrcProblem()
    for name in ("d", "n_cov", "nn_in", "nn_hidden", "nn_out", "rhs_kind", "n_scalars", "solver", "substeps",
                 "off_w1", "off_b1", "off_w2", "off_b2", "n_pm", "time_scale", "time_shift", "loss_kind",
                 "pred_length", "shoot_first_scored", "remove_ends", "preroll_eps", "n_series", "n_cols",
                 "n_models", "n_theta"):
        setattr(s, name, getattr(p, name))
    s.has_series_param = int(p.has_series_param)
    s.proc_inv_dt = int(p.proc_inv_dt)
    s.shoot_from_theta = int(p.shoot_from_theta)
    s.off_series_param = int(p.off_series_param)
    for k, o in enumerate(p.off_scalar):
        s.off_scalar[k] = int(o)
    keep = []
    for name in ("A", "B", "starts", "lengths", "model_of_series", "times", "data", "theta_base", "model_col0",
                 "model_series0", "obs_scale", "proc_scale", "shoot_weight", "reg_weight", "knot_off", "knot_t",
                 "knot_x", "dm_u", "dm_dudt", "skip_mask"):
        a = getattr(p, name)
        keep.append(a)
        setattr(s, name, _ptr(a))
    return s, keep


def max_threads() -> int:
    return int(lib().orc_max_threads())


def loss_grad(p, theta, want_grad: bool = True, threads: int = 1):
    """Returns (loss[n_models], grad[n_theta] or None)."""
    s, keep = _marshal(p)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    assert theta.shape[0] == p.n_theta, (theta.shape, p.n_theta)
    loss = np.zeros(p.n_models)
    grad = np.zeros(p.n_theta) if want_grad else None
    rc = lib().orc_loss_grad(C.byref(s), _ptr(theta), _ptr(loss), _ptr(grad), int(threads))
    if rc != 0:
        raise RuntimeError(f"orc_loss_grad failed rc={rc}")
    return loss, grad


def forward(p, theta, threads: int = 1):
    s, keep = _marshal(p)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    out = np.zeros((p.d, p.n_cols), order="F")
    rc = lib().orc_forward(C.byref(s), _ptr(theta), _ptr(out), int(threads))
    if rc != 0:
        raise RuntimeError(f"orc_forward failed rc={rc}")
    return out


def count_steps(p) -> int:
    s, keep = _marshal(p)
    return int(lib().orc_count_steps(C.byref(s)))
