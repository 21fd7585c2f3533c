"""Pins the oracle's rpbz / vdb arithmetic against the REFERENCE'S OWN object code.

oracle/_ref/libref_stats.so is src/ext/bcftools/bcftools-ext.c of the reference, compiled unmodified
from /root/reference by oracle/Makefile (two stub headers, no source copied). kf_erfc (htslib, not
vendored) is checked against scipy's erfc. calc_sor (src/plp_utils.c:375-387) is checked against an
independent numpy evaluation of the GATK formula.
"""
import ctypes as C
import math
import os

import numpy as np
import pytest

from oracle import pyoracle

REF_SO = os.path.join(os.path.dirname(pyoracle.__file__), "_ref", "libref_stats.so")


def _hist(rng, depth, spread):
    h = np.zeros(100, dtype=np.int32)
    if depth:
        idx = np.clip(rng.normal(50, spread, size=depth).astype(int), 0, 99)
        np.add.at(h, idx, 1)
    return h


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built (reference sources absent)")
def test_vdb_and_mwu_match_reference_object_code():
    ref = C.CDLL(REF_SO)
    ref.calc_vdb.argtypes = [C.POINTER(C.c_int), C.c_int]
    ref.calc_vdb.restype = C.c_double
    ref.calc_mwu_biasZ.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int]
    ref.calc_mwu_biasZ.restype = C.c_double
    L = pyoracle.lib()
    rng = np.random.default_rng(7)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))  # noqa: E731
    for trial in range(3000):
        da = int(rng.choice([0, 1, 2, 3, 5, 9, 10, 17, 40, 99, 150, 200, 450, 5000]))
        db = int(rng.choice([0, 1, 2, 3, 4, 7, 15, 16, 31, 50, 100, 199, 200, 201, 900]))
        a = _hist(rng, da, rng.uniform(1, 40))
        b = _hist(rng, db, rng.uniform(1, 40))
        v0, v1 = ref.calc_vdb(ip(b), 100), L.oracle_calc_vdb(ip(b), 100)
        assert v0 == v1 or (math.isinf(v0) and math.isinf(v1)), (trial, v0, v1)
        z0, z1 = ref.calc_mwu_biasZ(ip(a), ip(b), 100, 0, 1), L.oracle_calc_mwu_biasZ(ip(a), ip(b), 100, 0, 1)
        assert z0 == z1 or (math.isinf(z0) and math.isinf(z1)), (trial, z0, z1)


def test_kf_erfc_against_scipy():
    from scipy.special import erfc

    L = pyoracle.lib()
    xs = np.concatenate([np.linspace(-6, 6, 2001), np.array([-30.0, -10.0, 10.0, 26.0, 27.0, 40.0])])
    got = np.array([L.oracle_kf_erfc(float(x)) for x in xs])
    want = erfc(xs)
    ok = want > 1e-300
    assert np.allclose(got[ok], want[ok], rtol=2e-7, atol=0)
    assert np.all(got[~ok] < 1e-290)


def test_calc_sor_formula():
    L = pyoracle.lib()
    rng = np.random.default_rng(3)
    for _ in range(500):
        f_r, r_r, f_a, r_a = [int(x) for x in rng.integers(0, 300, size=4)]
        t00, t01, t11, t10 = f_r + 1.0, r_r + 1.0, f_a + 1.0, r_a + 1.0
        ratio = (t00 / t01) * (t11 / t10) + (t01 / t00) * (t10 / t11)
        want = math.log(ratio) + math.log(min(t00, t01) / max(t00, t01)) - math.log(min(t10, t11) / max(t10, t11))
        assert L.oracle_calc_sor(f_r, r_r, f_a, r_a) == pytest.approx(want, rel=1e-14, abs=1e-15)
