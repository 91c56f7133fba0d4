"""User-written right-hand sides (`derivs!` closures outside the catalog) on the CUDA hot path.

The reference lets the user write any `derivs!(du, u, p, t)` around a neural network (src/Models.jl:115-116, :148;
README.md:46-62; docs/src/examples.md); Julia compiles the closure and Zygote differentiates it.  A hand-written kernel cannot
run a Python closure, so the same freedom is offered one level down: the parametric part of the right-hand side is given as
EXPRESSIONS,

    rhs = expression_rhs(d=2, nn_out=1, scalars=("r", "m", "theta"),
                         du=["r*u1 - y1", "theta*y1 - m*u2"])                      # test/UDETests.jl:20-28

in the symbols u1..ud (states), x1..xk (covariates at the stage time), y1..ym (outputs of the network NN([u; X]) or
NN([u; X; a t + b]) with `time_input=(a, b)`), the named scalar parameters, and + - * / ** abs exp log sqrt tanh.  This
module differentiates them symbolically (SymPy), writes the `input / combine / combine_vjp` trait the lane-per-hidden-unit
kernel family is templated on (csrc/ude_kernel_body.inc: the same interface the catalog entries implement by hand), compiles
that ONE trait into the fused RK + loss + discrete-adjoint kernels with nvcc for sm_100a, and loads the result as a plugin of
libude_cuda.so: the plugin's static initialiser adds its kernels to the library's registry under a fresh `rhs_kind`, and from
there on `ude_create` / `ude_loss_grad` / `ude_adam` / `ude_forward` / the UKF treat it like any catalog entry.  The network
sees every state (the kernels return the input gradient straight into the state adjoint); FP64, or the optional FP32 mode.

The plugin cache (`csrc/user_rhs/`) is keyed by the generated source, the kernel geometry and the library headers.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
import threading
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

from . import problem as P
from . import build as _build

USER_KIND0 = 1000                      # rhs_kind values from here on are generated traits
_CACHE = os.path.join(_build.CSRC, "user_rhs")
_loaded: dict = {}                     # plugin path -> ctypes handle (kept alive: the registry points into it)
_by_kind: dict = {}                    # rhs_kind -> UserRhs
_lock = threading.RLock()              # one plugin build / load at a time per process (training.adam_multi_gpu: a thread per GPU)


def _sympy():
    import sympy
    return sympy


@dataclass(frozen=True)
class UserRhs:
    """What `derivs` may be besides a catalog entry (models.Rhs): carries the generated CUDA trait and a NumPy twin."""
    kind: int
    d: int
    n_cov: int
    nn_in: int
    nn_out: int
    scalar_names: tuple
    time_scale: float
    time_shift: float
    has_time_input: bool
    cuda_trait: str = field(repr=False, default="")
    du_numpy: Optional[Callable] = field(repr=False, default=None, compare=False)      # (u, X, y, sc) -> du, complex-step safe
    series_param: Optional[str] = None
    name: str = "USER"


def _printer():
    sp = _sympy()
    from sympy.printing.c import C99CodePrinter

    class Printer(C99CodePrinter):
        def _print_Pow(self, e):                     # small integer powers as products (pow() is slow and loses the sign logic)
            b, ex = e.as_base_exp()
            if ex.is_Integer and 2 <= int(ex) <= 4:
                s = self.parenthesize(b, 1000)
                return "(" + "*".join([s] * int(ex)) + ")"
            if ex.is_Integer and -4 <= int(ex) <= -1:
                s = self.parenthesize(b, 1000)
                return "(real(1)/(" + "*".join([s] * (-int(ex))) + "))"
            return super()._print_Pow(e)

        def _print_sign(self, e):                    # sgn1 of ude_kernel_body.inc: sign(0) = +1
            return "sgn1(" + self._print(e.args[0]) + ")"

        def _print_Float(self, e):
            return "real(" + repr(float(e)) + ")"

        def _print_Integer(self, e):
            return "real(" + str(int(e)) + ")"

        def _print_Rational(self, e):
            return "(real(%d)/real(%d))" % (e.p, e.q)

    return Printer()


def expression_rhs(d: int, nn_out: int, du: Sequence[str], scalars: Sequence[str] = (), n_cov: int = 0,
                   time_input: Optional[tuple] = None, series_param: Optional[str] = None, name: str = "USER") -> UserRhs:
    """Build a right-hand side from expressions (module docstring).  `du[i]` is d u_{i+1} / dt.  `series_param="r"` adds one
    trainable parameter PER SERIES (MultiCustomDerivatives' `p.r[i]`, docs/src/MultipleTimeSeries.md:51-55), visible in the
    expressions under that name."""
    sp = _sympy()
    if len(du) != d:
        raise ValueError("one expression per state")
    if not (1 <= d <= 8) or nn_out < 1:                               # UDE_MAX_D
        raise ValueError("d must be 1..8 and nn_out >= 1")
    if len(scalars) > 8:
        raise ValueError("at most 8 scalar parameters (UDE_MAX_SCALARS)")
    us = [sp.Symbol(f"u{i + 1}", real=True) for i in range(d)]
    xs = [sp.Symbol(f"x{i + 1}", real=True) for i in range(n_cov)]
    ys = [sp.Symbol(f"y{i + 1}", real=True) for i in range(nn_out)]
    ps = [sp.Symbol("p_" + str(n), real=True) for n in scalars]          # prefixed: a parameter may be called `c`, `t`, `y`, ...
    table = {s.name: s for s in us + xs + ys}
    for n, s in zip(scalars, ps):
        if str(n) in table or not str(n).isidentifier():
            raise ValueError(f"scalar parameter name {n!r} is not an identifier or shadows a state / covariate / network output")
        table[str(n)] = s
    sps = []
    if series_param is not None:
        if str(series_param) in table or not str(series_param).isidentifier():
            raise ValueError(f"per-series parameter name {series_param!r} is not an identifier or is already taken")
        sps = [sp.Symbol("ps_" + str(series_param), real=True)]
        table[str(series_param)] = sps[0]
    table.update(abs=sp.Abs, exp=sp.exp, log=sp.log, sqrt=sp.sqrt, tanh=sp.tanh)
    exprs = []
    for e in du:
        ex = sp.sympify(e, locals=table)
        unknown = ex.free_symbols - set(us + xs + ys + ps + sps)
        if unknown:
            raise ValueError(f"unknown symbols {sorted(map(str, unknown))} in {e!r}: use u1..u{d}, x1..x{n_cov}, y1..y{nn_out} and {list(scalars)}")
        exprs.append(ex)
    kb = [sp.Symbol(f"kb{i}", real=True) for i in range(d)]
    seed = sum(k * ex for k, ex in zip(kb, exprs))
    yb = [sp.diff(seed, y) for y in ys]
    ub = [sp.diff(seed, u) for u in us]
    gp = [sp.diff(seed, p) for p in ps]
    gs = [sp.diff(seed, q) for q in sps]
    need_y = any(y in ex.free_symbols for ex in yb + ub + gp + gs for y in ys)
    pr = _printer()

    def block(targets, values, op):
        """CSE'd assignments `target op value;` as C."""
        repl, red = sp.cse(list(values), symbols=sp.numbered_symbols("t_"), optimizations="basic")
        lines = [f"    const real {pr.doprint(s)} = {pr.doprint(v)};" for s, v in repl]
        lines += [f"    {t} {op} {pr.doprint(v)};" for t, v in zip(targets, red)]
        return "\n".join(lines)

    def binds(used):
        out = []
        for i, s in enumerate(us):
            if s in used:
                out.append(f"    const real u{i + 1} = U[{i}];")
        for i, s in enumerate(xs):
            if s in used:
                out.append(f"    const real x{i + 1} = X[{i}];")
        for i, s in enumerate(ys):
            if s in used:
                out.append(f"    const real y{i + 1} = y[{i}];")
        for i, s in enumerate(ps):
            if s in used:
                out.append(f"    const real {s.name} = c.sc[{i}];")
        for s in sps:
            if s in used:
                out.append(f"    const real {s.name} = c.series_param;")
        return "\n".join(out)

    free_f = set().union(*[e.free_symbols for e in exprs])
    free_b = set().union(*[e.free_symbols for e in yb + ub + gp + gs]) if (yb or ub or gp or gs) else set()
    din = d + n_cov + (1 if time_input is not None else 0)
    tin = f"    x[{d + n_cov}] = (real)(c.time_scale * t + c.time_shift);\n" if time_input is not None else ""
    kb_binds = "\n".join(f"    const real kb{i} = kb[{i}];" for i in range(d))
    trait = f"""struct RhsUser {{
  static constexpr int KIND = @KIND@, D = {d}, NX = {n_cov}, DIN = {din}, DOUT = {nn_out}, NS = {len(ps)};
  static constexpr bool NEED_Y = {"true" if need_y else "false"}, SERIES = {"true" if sps else "false"};
  __device__ __forceinline__ static void input(const real (&U)[D], const real* X, double t, const RhsCtx& c, real (&x)[DIN]) {{
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = U[j];
#pragma unroll
    for (int v = 0; v < NX; ++v) x[D + v] = X[v];
{tin}  }}
  __device__ __forceinline__ static void combine(const real (&U)[D], const real* X, const real (&y)[DOUT], const RhsCtx& c, real (&du)[D]) {{
{binds(free_f)}
{block([f"du[{i}]" for i in range(d)], exprs, "=")}
  }}
  __device__ __forceinline__ static void combine_vjp(const real (&U)[D], const real* X, const real (&y)[DOUT], const RhsCtx& c,
                                                     const real (&kb)[D], real (&yb)[DOUT], real (&ub)[D], real* gsc, real& gseries) {{
{kb_binds}
{binds(free_b)}
{block([f"yb[{i}]" for i in range(nn_out)] + [f"ub[{i}]" for i in range(d)] + [f"gsc[{i}]" for i in range(len(ps))] + ["gseries"] * len(sps),
       yb + ub + gp + gs, "@OP@")}
  }}
}};
"""
    # yb is assigned, ub and gsc accumulate
    fixed = []
    for ln in trait.split("\n"):
        if "@OP@" in ln:
            ln = ln.replace("@OP@", "=" if ln.strip().startswith("yb[") else "+=")
        fixed.append(ln)
    trait = "\n".join(fixed)
    kind = USER_KIND0 + int(hashlib.sha1(trait.encode()).hexdigest()[:6], 16)
    trait = trait.replace("@KIND@", str(kind)).replace("struct RhsUser {", f"struct RhsUser{kind} {{")      # one type name per right-hand side

    def _abs_cs(z):
        return z * (-1.0 if np.real(z) < 0 else 1.0)

    f_np = sp.lambdify([us, xs, ys, ps, sps], exprs, modules=[{"Abs": _abs_cs, "sign": lambda z: -1.0 if np.real(z) < 0 else 1.0}, "numpy"])

    def du_numpy(u, X, y, sc, series_value=None):
        return np.array(f_np(list(u), list(X), list(y), list(sc), [series_value] if sps else []))

    ts, tsh = (float(time_input[0]), float(time_input[1])) if time_input is not None else (1.0, 0.0)
    rhs = UserRhs(kind, d, n_cov, din, nn_out, tuple(str(s) for s in scalars), ts, tsh, time_input is not None, trait, du_numpy,
                  str(series_param) if series_param is not None else None, name)
    _by_kind[kind] = rhs
    return rhs


def lookup(kind: int) -> Optional[UserRhs]:
    return _by_kind.get(int(kind))


def geometry(hidden: int) -> tuple:
    """(lanes per segment G, hidden units per lane HPL, resident blocks per SM asked of the compiler) for a hidden width --
    the choices the catalog instances make (csrc/ude_inst_*.cu)."""
    if hidden <= 10:
        return 2, (hidden + 1) // 2, 2
    if hidden <= 12:
        return 4, 3, 4
    if hidden <= 32:
        return 8, (hidden + 7) // 8, 3
    if hidden <= 64:
        return 16, (hidden + 15) // 16, 3
    if hidden <= 128:
        return 32, (hidden + 31) // 32, 2
    raise ValueError("hidden width above 128")


def _lib_stamp() -> str:
    """changes whenever the headers the plugin is compiled against change (stale plugins must not be loaded)"""
    h = hashlib.sha1()
    for f in sorted(os.listdir(_build.CSRC)):
        if f.endswith((".cuh", ".h", ".inc")):
            with open(os.path.join(_build.CSRC, f), "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()[:10]


def _ensure_kernels(rhs: UserRhs, hidden: int, solver: int, verbose: bool = False, dtype: int = 0) -> str:
    """Compile (once) and load the plugin that holds this trait's kernels for one hidden width and solver: the three loss
    families, each with and without the reverse sweep.  dtype = 1: the optional FP32 mode (the same trait text compiled in the
    float copy of the kernel family).  Returns the plugin path."""
    G, HPL, minb = geometry(hidden)
    ns, tns = ("F64", "f64") if dtype == 0 else ("F32", "f32")
    tag = hashlib.sha1((rhs.cuda_trait + f"|{ns}|{G}|{HPL}|{minb}|{solver}|" + _lib_stamp()).encode()).hexdigest()[:16]
    so = os.path.join(_CACHE, f"ude_user_{tag}.so")
    if so in _loaded:
        return so
    if not os.path.exists(so):
        os.makedirs(_CACHE, exist_ok=True)
        src = os.path.join(_CACHE, f"ude_user_{tag}.cu")
        tmp = f".{os.getpid()}.tmp"                   # several ranks may build the same plugin at once: private files, atomic renames
        label = ("f32_" if dtype else "") + f"user{rhs.kind}_g{G}h{HPL}"
        regs = "\n".join(f'  register_one<{ns}, RhsUser{rhs.kind}, {G}, {HPL}, {minb}, {solver}, {lk}, false>("{label}", 0);'
                         for lk in (("LK_STATE",) if solver == P.EULER else ("LK_STATE", "LK_TRAJ", "LK_DM")))
        with open(src + tmp, "w") as fh:
            fh.write("// generated by universaldiffeq.jl_b200/user_rhs.py from the expressions of a user right-hand side -- do not edit\n"
                     f'#include "ude_inst.cuh"\nnamespace ude {{\nnamespace {tns} {{\n' + rhs.cuda_trait + f"}}  // namespace {tns}\n}}  // namespace ude\n"
                     f"namespace {{\nusing namespace ude;\nusing ude::{tns}::RhsUser{rhs.kind};\nRegisterKernels reg([] {{\n" + regs + "\n});\n}  // namespace\n")
        os.replace(src + tmp, src)
        if not os.path.exists(_build.SO):           # never rebuild here: the library may already be loaded in this process
            raise RuntimeError("libude_cuda.so is not built (python -c 'import __graft_entry__ as g; g.build()')")
        cmd = [_build.NVCC, *_build.FLAGS, "-I", _build.CSRC, "-shared", "-o", so + tmp, src, "-L", _build.CSRC, "-lude_cuda", "-lcudart",
               "-Xlinker", "-rpath", "-Xlinker", _build.CSRC]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for the generated right-hand side ({src}):\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr:
            print(r.stderr)
        os.replace(so + tmp, so)
    _loaded[so] = ctypes.CDLL(so, mode=ctypes.RTLD_GLOBAL)
    return so


# ---- catalog right-hand sides at shapes without a pre-compiled instance ------------------------------------------------------
def _catalog_trait(p) -> Optional[str]:
    """C++ name of the hand-written trait (csrc/ude_kernel_body.inc) for a catalog problem"""
    k, d, nx, do = p.rhs_kind, p.d, p.n_cov, p.nn_out
    if k == P.RHS_NODE and p.nn_in == d + nx and do == d:
        return f"RhsNode<{d}, {nx}>"
    if k == P.RHS_NODE_TIME and p.nn_in == d + nx + 1:
        return f"RhsNodeTime<{d}, {nx}, {do}>"
    if k == P.RHS_LV_UDE and d == 2 and nx in (0, 1):
        return f"RhsLvUde<{nx}>"
    if k == P.RHS_LV_UDE_ABS and d == 2 and nx == 0:
        return "RhsLvUdeAbs"
    if k == P.RHS_FISHERY and d == 2 and nx == 1:
        return "RhsFishery"
    if k == P.RHS_NN_DECAY and nx == 0:
        return f"RhsNnDecay<{d}>"
    if k == P.RHS_SERIES_GROWTH and nx == 0:
        return f"RhsSeriesGrowth<{d}>"
    return None


def _ensure_catalog_kernels(p, dtype: int = 0, verbose: bool = False) -> str:
    """The library ships the catalog right-hand sides at the shapes of the BASELINE configurations, the reference's tests and
    docs (hidden width <= 12 for the parametric UDEs, the NODE shapes of csrc/ude_inst_*.cu).  Any other hidden width up to 128,
    with one or many models per handle, is compiled here on first use -- the same hand-written trait, instantiated for the
    geometry `geometry(hidden)` picks -- and registered like a user right-hand side.  Raises ValueError if the problem is not a
    catalog shape."""
    trait = _catalog_trait(p)
    if trait is None:
        raise ValueError("not a catalog right-hand side")
    G, HPL, minb = geometry(p.nn_hidden)
    ns, tns = ("F64", "f64") if dtype == 0 else ("F32", "f32")
    tag = hashlib.sha1(f"{trait}|{ns}|{G}|{HPL}|{minb}|{p.solver}|{_lib_stamp()}".encode()).hexdigest()[:16]
    so = os.path.join(_CACHE, f"ude_cat_{tag}.so")
    if so in _loaded:
        return so
    if not os.path.exists(so):
        os.makedirs(_CACHE, exist_ok=True)
        src = os.path.join(_CACHE, f"ude_cat_{tag}.cu")
        tmp = f".{os.getpid()}.tmp"
        label = "".join(ch if ch.isalnum() else "_" for ch in trait).strip("_").lower() + f"_{tns}_g{G}h{HPL}_jit"
        regs = "\n".join(f'  register_one<{ns}, {tns}::{trait}, {G}, {HPL}, {minb}, {p.solver}, {lk}, false>("{label}", 5);'
                         for lk in (("LK_STATE",) if p.solver == P.EULER else ("LK_STATE", "LK_TRAJ", "LK_DM")))
        with open(src + tmp, "w") as fh:
            fh.write("// generated by universaldiffeq.jl_b200/user_rhs.py: a catalog right-hand side at a shape without a pre-compiled instance\n"
                     '#include "ude_inst.cuh"\nnamespace {\nusing namespace ude;\nRegisterKernels reg([] {\n' + regs + "\n});\n}  // namespace\n")
        os.replace(src + tmp, src)
        if not os.path.exists(_build.SO):           # never rebuild here: the library may already be loaded in this process
            raise RuntimeError("libude_cuda.so is not built (python -c 'import __graft_entry__ as g; g.build()')")
        cmd = [_build.NVCC, *_build.FLAGS, "-I", _build.CSRC, "-shared", "-o", so + tmp, src, "-L", _build.CSRC, "-lude_cuda", "-lcudart",
               "-Xlinker", "-rpath", "-Xlinker", _build.CSRC]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr:
            print(r.stderr)
        os.replace(so + tmp, so)
    _loaded[so] = ctypes.CDLL(so, mode=ctypes.RTLD_GLOBAL)
    return so


def ensure_kernels(rhs: UserRhs, hidden: int, solver: int, verbose: bool = False, dtype: int = 0) -> str:
    with _lock:
        return _ensure_kernels(rhs, hidden, solver, verbose, dtype)


def ensure_catalog_kernels(p, dtype: int = 0, verbose: bool = False) -> str:
    with _lock:
        return _ensure_catalog_kernels(p, dtype, verbose)


ensure_kernels.__doc__ = _ensure_kernels.__doc__
ensure_catalog_kernels.__doc__ = _ensure_catalog_kernels.__doc__
