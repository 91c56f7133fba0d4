"""The C-ABI library loads on a CPU-only box and exports every symbol include/ude_cuda.h declares; without a GPU the
compute entry points fail loudly (there is no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "ude_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ude_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(ude):
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(ude.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(ude._capi.EXPORTS) == names


def test_desc_layout_matches_header(ude):
    # sizeof(ude_desc) as the C compiler sees it == the ctypes mirror (ude_create rejects any other size)
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include "ude_cuda.h"\nint main(){printf("%zu %d", sizeof(ude_desc), UDE_ABI_VERSION);return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c"); exe = os.path.join(td, "s")
        open(c, "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        size, ver = subprocess.check_output([exe]).decode().split()
    assert int(size) == ctypes.sizeof(ude._capi.UdeDesc)
    assert int(ver) == ude._capi.ABI_VERSION


def test_no_cpu_fallback(ude):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert ude.device_count() == 0
    import cases
    prob, theta = cases.make_case("node_d2", "cond_lik", 0)
    with pytest.raises(ude.UdeError) as ei:
        ude.Engine(prob)
    assert ei.value.code == -4          # UDE_ERR_NO_DEVICE
    with pytest.raises(ude.UdeError):
        ude.measure_fp64_peak(0)


def test_product_path_never_imports_oracle():
    pkg = os.path.join(ROOT, "universaldiffeq.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "ude_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_julia_desc_mirror_matches_the_header():
    """julia/UDECuda.jl cannot be executed here (no Julia): at least its `Desc` struct must list the fields of `ude_desc` in the
    header's order with matching widths, and every function it ccalls must be an exported symbol."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "ude_cuda.h")).read()
    body = hdr[hdr.index("typedef struct ude_desc {"):hdr.index("} ude_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = []
    for ty, names in re.findall(r"\b(int32_t|int64_t|double)\s+([^;]+);", body):
        for nm in names.split(","):
            nm = nm.strip()
            m = re.match(r"(\w+)\[(.+)\]", nm)
            n = 1
            if m:
                nm, dim = m.group(1), m.group(2)
                n = eval(dim.replace("UDE_MAX_SCALARS", "8").replace("UDE_MAX_D", "8"))
            c_fields.append((nm, {"int32_t": "Int32", "int64_t": "Int64", "double": "Float64"}[ty], n))
    jl = open(os.path.join(root, "julia", "UDECuda.jl")).read()
    sb = jl[jl.index("struct Desc"):]
    sb = sb[:sb.index("\nend")]
    j_fields = []
    for nm, ty in re.findall(r"(\w+)::((?:NTuple\{\d+,\s*\w+\})|\w+)", sb):
        m = re.match(r"NTuple\{(\d+),\s*(\w+)\}", ty)
        j_fields.append((nm, m.group(2), int(m.group(1))) if m else (nm, ty, 1))
    assert j_fields == c_fields, (j_fields, c_fields)
    called = set(re.findall(r"ccall\(\(:(\w+),", jl))
    assert called and called <= set(_declared()), called - set(_declared())
