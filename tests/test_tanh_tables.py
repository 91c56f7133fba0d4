"""The 2^(j/n) tables behind the kernels' tanh (csrc/ude_math.cuh) are correctly rounded, and the arithmetic of the shipped
variant -- re-stated here in NumPy float64 without FMA -- stays at the 1e-15 level against math.tanh, saturates for large
arguments and uses the biased-table exponent patch consistently.  CPU only: the device versions are measured by
tools/tanh_check.cu (profiles/tanh_check_r02*.json) and exercised by every GPU parity test."""
import math
import os
import re
from decimal import Decimal, getcontext

import numpy as np

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "universaldiffeq.jl_b200", "csrc")


def _table(name):
    txt = open(os.path.join(CSRC, name)).read()
    return [float.fromhex(x) for x in re.findall(r"0x1\.[0-9a-f]+p[+-]\d+", txt)]


def test_exp2_tables_are_correctly_rounded():
    getcontext().prec = 60
    ln2 = Decimal(2).ln()
    for name, n in (("ude_exp2_table.inc", 64), ("ude_exp2_table256.inc", 256)):
        tab = _table(name)
        assert len(tab) == n
        for j, v in enumerate(tab):
            assert v == float((Decimal(j) / Decimal(n) * ln2).exp()), (name, j)
    assert _table("ude_exp2_table256.inc")[::4] == _table("ude_exp2_table.inc")


def _tanh_variant(x, n_tab, shift_bits, coeffs):
    """ude_tanh_fast3_vec (n_tab = 64) / ude_tanh_fast4_vec (n_tab = 256) in NumPy: signed argument, clamp, k = round(x 2n/ln2),
    biased table entry + (N << shift_bits), E = exp(2 rh) by Horner, w + 1 = s E + 1, tanh = 1 - 2/(w + 1)."""
    tab = np.array(_table("ude_exp2_table.inc" if n_tab == 64 else "ude_exp2_table256.inc"))
    hi = (tab.view(np.int64) >> 32).astype(np.int64) - (np.arange(n_tab, dtype=np.int64) << shift_bits)      # biased high words
    lo = tab.view(np.int64) & 0xFFFFFFFF
    xc = np.clip(x, -20.0, 20.0)
    inv, ch = 2 * n_tab / math.log(2.0), math.log(2.0) / (2 * n_tab)
    k = np.rint(xc * inv)
    N = k.astype(np.int64)
    rh = xc - k * ch
    q = np.full_like(xc, coeffs[0])
    for c in coeffs[1:]:
        q = q * rh + c
    j = N & (n_tab - 1)
    s_bits = ((hi[j] + (N << shift_bits)) << 32) | lo[j]
    s = s_bits.view(np.float64)
    assert np.allclose(s, np.exp2(N / n_tab), rtol=1e-15)          # the patched entry is 2^(N/n): the bias cancels exactly
    return 1.0 - 2.0 / (s * q + 1.0)


def test_tanh_variants_restated():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-22, 22, 200_000), rng.uniform(-0.05, 0.05, 50_000), rng.normal(size=50_000) * 1e3,
                        np.array([0.0, -0.0, 20.0, -20.0, 1e300, -1e300, np.inf, -np.inf])])
    ref = np.tanh(x)
    for n_tab, sh, co in ((64, 14, (4 / 15, 2 / 3, 4 / 3, 2.0, 2.0, 1.0)), (256, 12, (2 / 3, 4 / 3, 2.0, 2.0, 1.0))):
        y = _tanh_variant(x, n_tab, sh, co)
        assert np.max(np.abs(y - ref)) < 1.5e-15, (n_tab, np.max(np.abs(y - ref)))
        assert np.all(y[np.abs(x) > 20.0] == np.sign(x[np.abs(x) > 20.0]))
