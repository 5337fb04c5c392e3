"""CPU model of the fixed-point source coordinates of csrc/edf_fx.cuh (no GPU): the bit manipulations the kernels
apply to T = c + 1.5 * 2^29 -- floor from bits [23, 55), fraction from the low 23 bits, validity from the exponent
field, the strict range tests of odd and even orders -- restated with NumPy integer views and checked against plain
floating-point arithmetic.  The constants are read from the header, so the model and the kernels cannot drift apart."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "elasticdeform_b200", "csrc", "edf_fx.cuh")).read()


def _const(name):
    m = re.search(r"#define\s+%s\s+([0-9a-fA-Fx.eE+-]+)u?" % name, HDR)
    assert m, name
    return m.group(1)


FBITS = int(_const("EDF_PP_FBITS"))
MAGIC = float(_const("EDF_PP_MAGIC"))
HI0 = int(_const("EDF_PP_HI0"), 16)
FLBIAS = int(_const("EDF_PP_FLBIAS"), 16)


def split(T):
    """(floor, fraction bits, valid) exactly as edf_pipe_coords / edf_fx_code compute them."""
    bits = np.asarray(T, dtype=np.float64).view(np.uint64)
    lo = (bits & np.uint64(0xFFFFFFFF)).astype(np.uint64)
    hi = (bits >> np.uint64(32)).astype(np.uint64)
    fs = ((bits >> np.uint64(FBITS)) & np.uint64(0xFFFFFFFF))                       # __funnelshift_r(lo, hi, 23)
    fl = ((fs - np.uint64(FLBIAS)) & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.int32)
    gq = (lo & np.uint64((1 << FBITS) - 1)).astype(np.int64)
    valid = ((hi - np.uint64(HI0)) & np.uint64(0xFFFFFFFF)) < np.uint64(0x00100000)
    return fl, gq, valid


def test_constants():
    assert FBITS == 23 and MAGIC == 1.5 * 2.0 ** 29
    bits = np.float64(MAGIC).view(np.uint64)
    assert int(bits >> np.uint64(32)) & 0xFFF00000 == HI0                            # exponent field of 2^29
    assert int((bits >> np.uint64(FBITS)) & np.uint64(0xFFFFFFFF)) == FLBIAS         # bits [23,55) of the magic number


def test_floor_and_fraction_from_the_bits():
    rng = np.random.default_rng(0)
    c = np.concatenate([rng.uniform(-3000.0, 70000.0, 200000), rng.uniform(-4.0, 4.0, 100000),
                        np.arange(-64, 64, dtype=np.float64), np.arange(-64, 64) + 0.5,
                        np.arange(-64, 64) + 2.0 ** -23, np.arange(-64, 64) - 2.0 ** -23])
    T = c + MAGIC                                                                    # the last FMA of the Horner form
    fl, gq, valid = split(T)
    assert valid.all()
    q = np.round(c * 2.0 ** FBITS)                                                   # c on the 2^-23 grid (ties to even, as the FMA rounds)
    np.testing.assert_array_equal(fl, np.floor(q / 2.0 ** FBITS).astype(np.int64))
    np.testing.assert_array_equal(gq, (q - np.floor(q / 2.0 ** FBITS) * 2.0 ** FBITS).astype(np.int64))
    # the fraction as the kernels build it: the 23 bits placed in the mantissa of a float in [1, 2), minus 1.5
    e = ((gq.astype(np.uint32) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.5)).astype(np.float64)
    np.testing.assert_array_equal(e, gq / 2.0 ** FBITS - 0.5)                        # exact, in [-0.5, 0.5)
    assert np.abs((fl + gq / 2.0 ** FBITS) - c).max() <= 2.0 ** -24 * (1 + 1e-9)     # quantisation: half a grid step


def test_out_of_range_and_nonfinite_coordinates_are_flagged():
    bad = np.array([2.0 ** 28, -2.0 ** 28 - 1, 1e12, -1e12, np.inf, -np.inf, np.nan])
    with np.errstate(invalid="ignore"):
        _, _, valid = split(bad + MAGIC)
    assert not valid.any()
    ok = np.array([2.0 ** 28 - 1, -2.0 ** 28 + 1, 0.0, -0.0])
    assert split(ok + MAGIC)[2].all()


def test_strict_range_tests():
    """Odd orders: 0 <= floor(c) <= len - 2; even orders (T holds c + 0.5): 1 <= floor(2c + 1) <= 2 len - 2.  Both are
    the reference's `0 <= c <= len - 1` except ON the limits, where the voxel is redone exactly."""
    rng = np.random.default_rng(1)
    n = 57
    c = rng.uniform(-3.0, n + 2.0, 400000)
    c = c[np.abs(c - np.round(c)) > 1e-6]                                            # away from the integers (the limits)
    inr_ref = (c >= 0) & (c <= n - 1)
    fl, _, _ = split(c + MAGIC)
    np.testing.assert_array_equal(fl.astype(np.uint32) <= np.uint32(n - 2), inr_ref)
    c2 = c[np.abs(2 * c - np.round(2 * c)) > 1e-6]                                   # even orders: away from the half-integers too
    T = (c2 + 0.5) + MAGIC
    bits = T.view(np.uint64)
    h = (((bits >> np.uint64(FBITS - 1)) & np.uint64(0xFFFFFFFF)) - np.uint64(2 * FLBIAS) - np.uint64(1)) & np.uint64(0xFFFFFFFF)
    np.testing.assert_array_equal(h.astype(np.uint32) <= np.uint32(2 * n - 3), (c2 >= 0) & (c2 <= n - 1))
    fl2, _, _ = split(T)
    np.testing.assert_array_equal(fl2, np.floor(np.round((c2 + 0.5) * 2.0 ** FBITS) / 2.0 ** FBITS).astype(np.int64))
