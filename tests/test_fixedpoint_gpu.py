"""Parity of the kernels that take their coordinates from the folded fixed-point polynomial (csrc/edf_poly.cuh:
edf_poly_fold / edf_pipe_coords) against the oracle -- `-m gpu`.  At orders 0 / 1 in 'constant' mode that is the
direct forward kernel ``poly3d_f32_direct``; the cases make sure it is the kernel that ran and cover what is special
about the formulation: tiles that are not full, voxels at every border of the volume, crop offsets, a 3-D affine map
folded into the polynomial, identity / zero displacement (exact integer coordinates, all on thresholds), steep
fields, non-finite control points.  Orders 2 / 3 run the same cases through whatever kernel the library picks."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def edf():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs the GPU box")
    import elasticdeform_b200
    return elasticdeform_b200


def _impl():
    return "ref" if O.ref_available() else "port"


def _check(edf, X, D, order, expect_kernel="auto", **kw):
    """The host-side steep-field hint (a performance hint that routes steep fields to the round-1 kernels) is
    cleared, so that the kernel choice does not depend on the field."""
    import importlib
    from elasticdeform_b200 import _lib
    dg = importlib.import_module("elasticdeform_b200.deform_grid")
    saved = dg._steep_hint
    dg._steep_hint = lambda *a, **k: 0
    try:
        y = edf.deform_grid(X, D, order=order, prefilter=False, **kw)
    finally:
        dg._steep_hint = saved
    k = _lib.last_kernel()
    if expect_kernel == "auto":
        expect_kernel = "poly3d_f32_direct" if order <= 1 else None
    if expect_kernel:
        assert k == expect_kernel, (k, order, kw.keys())
    yr = O.deform_grid(X, D, order=order, prefilter=False, impl=_impl(), **kw)
    assert y.shape == yr.shape and y.dtype == yr.dtype
    if order == 0:
        np.testing.assert_array_equal(y, yr)
    else:
        np.testing.assert_allclose(y, yr, rtol=0, atol=1e-5)
    return y


@pytest.mark.parametrize("order", [0, 1, 2, 3])
@pytest.mark.parametrize("shape,points,sigma", [
    ((40, 72, 96), (5, 5, 5), 3.0),
    ((33, 41, 64), (3, 4, 5), 6.0),        # partial tiles along every axis
    ((70, 50, 132), (5, 5, 5), 15.0),      # many out-of-range voxels, voxels next to every border
])
def test_fixedpoint_forward(edf, order, shape, points, sigma):
    rng = np.random.default_rng(1000 + order)
    X = rng.random(shape, dtype=np.float32)
    D = rng.standard_normal((3,) + points) * sigma
    _check(edf, X, D, order)


@pytest.mark.parametrize("order", [0, 1, 2, 3])
def test_fixedpoint_identity_and_tiny_fields(edf, order):
    """Zero displacement: every coordinate is an integer (on a threshold); the result is the input itself (the
    last plane / row / column included: coordinate == len - 1 is in range, deform.c:84-86)."""
    rng = np.random.default_rng(2000 + order)
    X = rng.random((24, 40, 80), dtype=np.float32)
    y = _check(edf, X, np.zeros((3, 3, 3, 3)), order)
    if order <= 1:
        np.testing.assert_array_equal(y, X)
    _check(edf, X, rng.standard_normal((3, 3, 3, 3)) * 1e-9, order)
    _check(edf, X, rng.standard_normal((3, 3, 3, 3)) * 1e-3, order)


@pytest.mark.parametrize("order", [0, 1, 3])
def test_fixedpoint_crop_and_affine(edf, order):
    rng = np.random.default_rng(3000 + order)
    X = rng.random((48, 80, 112), dtype=np.float32)
    D = rng.standard_normal((3, 4, 5, 6)) * 4.0
    crop = (slice(5, 41), slice(16, 70), slice(30, 111))
    _check(edf, X, D, order, crop=crop)
    th = np.deg2rad(12.0)
    A = np.array([[1.1, 0.0, 0.0, -2.0],
                  [0.0, np.cos(th), -np.sin(th), 9.0],
                  [0.0, np.sin(th), np.cos(th), -7.5]])
    _check(edf, X, D, order, affine=A)
    _check(edf, X, D, order, affine=A, crop=crop)
    _check(edf, X, np.zeros((3, 3, 3, 3)), order, affine=A)           # affine coordinates without a displacement


@pytest.mark.parametrize("order", [1, 3])
def test_fixedpoint_steep_field(edf, order):
    """Boxes that outgrow a stage: shorter chunks, then the single-voxel routine (EDF_FLAG_STEEP would route the call
    to the round-1 kernels; the hint is cleared here to keep the pipelined kernel under test)."""
    rng = np.random.default_rng(4000 + order)
    X = rng.random((40, 64, 96), dtype=np.float32)
    D = rng.standard_normal((3, 6, 6, 6)) * 20.0
    _check(edf, X, D, order)


def test_fixedpoint_nonfinite_control_points(edf):
    rng = np.random.default_rng(5000)
    X = rng.random((24, 40, 64), dtype=np.float32)
    D = rng.standard_normal((3, 4, 4, 4)) * 2.0
    D[1, 2, 2, 2] = np.nan
    D[0, 0, 1, 3] = 1e300
    import importlib
    dg = importlib.import_module("elasticdeform_b200.deform_grid")
    saved = dg._steep_hint
    dg._steep_hint = lambda *a, **k: 0
    try:
        for order in (0, 3):
            y = edf.deform_grid(X, D, order=order, prefilter=False)
            yr = O.deform_grid(X, D, order=order, prefilter=False, impl=_impl())
            np.testing.assert_array_equal(np.isnan(y), np.isnan(yr))
            np.testing.assert_allclose(np.nan_to_num(y), np.nan_to_num(yr), rtol=0, atol=1e-5)
    finally:
        dg._steep_hint = saved


def test_fixedpoint_headline_slab(edf):
    """256^3, order 3, sigma 8 (the bench workload): 12 output planes through the crop offset against the oracle, and
    the same planes of the full-volume call (the crop call and the full call must agree bit for bit)."""
    from elasticdeform_b200 import _lib
    rng = np.random.default_rng(0)
    X = rng.random((256, 256, 256), dtype=np.float32)
    D = rng.standard_normal((3, 5, 5, 5)) * 8.0
    full = edf.deform_grid(X, D, order=3, prefilter=False)
    for z0 in (0, 122, 244):
        crop = (slice(z0, z0 + 12), slice(None), slice(None))
        part = edf.deform_grid(X, D, order=3, prefilter=False, crop=crop)
        ref = O.deform_grid(X, D, order=3, prefilter=False, crop=crop, impl=_impl())
        np.testing.assert_allclose(part, ref, rtol=0, atol=1e-5)
        np.testing.assert_allclose(full[crop], ref, rtol=0, atol=1e-5)


@pytest.mark.parametrize("dtype", [np.int32, np.uint32])
def test_fixedpoint_label_volume_bit_copy(edf, dtype):
    """4-byte label volumes at order 0 take the same direct kernel (the output is a bit copy of the selected voxel,
    or the converted cval): BASELINE config 3's int32 label."""
    from elasticdeform_b200 import _lib
    rng = np.random.default_rng(6000)
    L = rng.integers(0, 2 ** 31 - 1, (40, 56, 72)).astype(dtype)
    D = rng.standard_normal((3, 4, 4, 4)) * 5.0
    for kw in (dict(), dict(cval=7), dict(crop=(slice(3, 30), slice(None), slice(10, 60)))):
        y = edf.deform_grid(L, D, order=0, prefilter=False, **kw)
        assert _lib.last_kernel() == "poly3d_f32_direct", _lib.last_kernel()
        yr = O.deform_grid(L, D, order=0, prefilter=False, impl=_impl(), **kw)
        assert y.dtype == yr.dtype
        np.testing.assert_array_equal(y, yr)


@pytest.mark.parametrize("order", [0, 1])
def test_fixedpoint_channels_share_coordinates(edf, order):
    """One array of channels sharing a displacement (axis=(1, 2, 3), BASELINE config 5): one coordinate pass per
    voxel, the channel loop inside the kernel."""
    from elasticdeform_b200 import _lib
    rng = np.random.default_rng(6100 + order)
    X = rng.random((5, 40, 48, 72), dtype=np.float32)
    D = rng.standard_normal((3, 5, 5, 5)) * 4.0
    for kw in (dict(), dict(crop=(slice(4, 36), slice(8, 40), slice(0, 64)))):
        y = edf.deform_grid(X, D, order=order, axis=(1, 2, 3), prefilter=False, **kw)
        assert _lib.last_kernel() == "poly3d_f32_direct", _lib.last_kernel()
        yr = O.deform_grid(X, D, order=order, axis=(1, 2, 3), prefilter=False, impl=_impl(), **kw)
        if order == 0:
            np.testing.assert_array_equal(y, yr)
        else:
            np.testing.assert_allclose(y, yr, rtol=0, atol=1e-5)
    # a strided view (every other channel): the step stride is not the dense one
    Xv = X[::2]
    y = edf.deform_grid(Xv, D, order=order, axis=(1, 2, 3), prefilter=False)
    yr = O.deform_grid(np.ascontiguousarray(Xv), D, order=order, axis=(1, 2, 3), prefilter=False, impl=_impl())
    np.testing.assert_allclose(y, yr, rtol=0, atol=1e-5)
