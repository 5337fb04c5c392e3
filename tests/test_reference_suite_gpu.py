"""The reference's own acceptance suite, run against the CUDA package (`-m gpu`).

A seeded restatement of the cases of the reference's tests/test_deform_grid.py (cited below as
``ref:LINE``): the package under test is imported under the alias ``elasticdeform`` (and
``elasticdeform.torch``), exactly as a user switching over would, and every case of the 17
non-TensorFlow tests runs through it: the SciPy differential (ref:355-365 against the
``map_coordinates`` restatement ref:36-72), finite-difference gradients (ref:325-353),
crop == rotate/zoom-then-crop (ref:121-133), list == item-by-item (ref:294-323) and the torch
twins (ref:470-568).  /root/reference does not exist on the GPU box, so nothing is read from it;
the only checker here is SciPy + NumPy, as in the reference.  The TensorFlow wrapper is out of
scope (no TensorFlow in the image; the reference skips those two tests as well).
"""
import itertools
import sys

import numpy as np
import pytest
import scipy
import scipy.ndimage
from packaging import version

pytestmark = pytest.mark.gpu

ALL_MODES = ("nearest", "wrap", "reflect", "mirror", "constant")
# SciPy >= 1.6 changed 'reflect' and 'nearest' in map_coordinates; the reference skips the SciPy
# comparison for those (ref:30-33, ref:98-100) -- they are pinned by the compiled reference in test_parity_gpu.py
MODERN_SCIPY = version.parse(scipy.__version__) > version.parse("1.5.4")


@pytest.fixture(scope="module")
def ed():
    """``import elasticdeform`` -> the CUDA package (module alias, as a drop-in user would install it)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("these tests need the GPU box")
    import elasticdeform_b200
    import elasticdeform_b200.torch as etorch
    saved = {k: sys.modules.get(k) for k in ("elasticdeform", "elasticdeform.torch")}
    sys.modules["elasticdeform"] = elasticdeform_b200
    sys.modules["elasticdeform.torch"] = etorch
    import elasticdeform
    assert elasticdeform is elasticdeform_b200
    yield elasticdeform
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def scipy_deform(X, displacement, order=3, mode="constant", cval=0.0, crop=None, prefilter=True, axis=None):
    """What deform_grid computes, said with scipy.ndimage.map_coordinates (the construction of ref:36-72):
    the control grid is sampled at linspace(0, P-1, n) per deformed axis with a cubic spline, added to the
    voxel coordinates, and every slice over the non-deformed axes is resampled there."""
    axis = tuple(range(X.ndim)) if axis is None else ((axis,) if isinstance(axis, int) else tuple(axis))
    npts = displacement.shape[1:]
    where = np.meshgrid(*[np.arange(X.shape[a]) for a in axis], indexing="ij")
    ctrl = np.meshgrid(*[np.linspace(0, p - 1, X.shape[a]) for a, p in zip(axis, npts)], indexing="ij")
    full_crop = [slice(None)] * X.ndim
    if crop is not None:
        where = [w[crop] for w in where]
        ctrl = [c[crop] for c in ctrl]
        for a, c in zip(axis, crop):
            full_crop[a] = c
    full_crop = tuple(full_crop)
    where = [w + scipy.ndimage.map_coordinates(displacement[i], ctrl, order=3) for i, w in enumerate(where)]
    out = np.zeros(X[full_crop].shape, dtype=X.dtype)
    others = [[slice(None)] if a in axis else range(X.shape[a]) for a in range(X.ndim)]
    for idx in itertools.product(*others):
        scipy.ndimage.map_coordinates(X[idx], where, output=out[idx], order=order, cval=cval, mode=mode,
                                      prefilter=prefilter)
    return out


def compare_with_scipy(ed, rng, shape, points, order=3, sigma=25, crop=None, mode="constant", axis=None):
    """ref:355-365"""
    D = rng.standard_normal((len(shape) if axis is None else len(axis),) + tuple(points)) * sigma
    X = rng.random(shape)
    want = scipy_deform(X, D, order=order, crop=crop, mode=mode, axis=axis)
    got = ed.deform_grid(X, D, order=order, crop=crop, mode=mode, axis=axis)
    np.testing.assert_allclose(want, got, rtol=1e-05, atol=1e-08)


def check_gradient_numerically(rng, X, fn, grad_fn, eps=1e-4, n_tests=10):
    """ref:325-353: d/dX_i of sum(fn(X) * R) by one-sided differences against grad_fn(R)."""
    out_shape = fn(X).shape
    for _ in range(n_tests):
        R = rng.random(out_shape) + 0.5
        f0 = np.sum(fn(X) * R)
        num = np.zeros_like(X)
        Xp = X.copy()
        for i in range(X.size):
            Xp[:] = X
            Xp.flat[i] += eps
            num.flat[i] = (np.sum(fn(Xp) * R) - f0) / eps
        np.testing.assert_allclose(num, grad_fn(R, X), rtol=1e-05, atol=1e-08)


def test_random(ed):
    """ref:89-94"""
    rng = np.random.default_rng(100)
    for points in (3, (3, 5)):
        for shape in ((100, 100), (100, 75)):
            for order in (0, 1, 2, 3, 4):
                Y = ed.deform_random_grid(rng.random(shape), points=points, order=order)
                assert Y.shape == shape


def test_basic_2d(ed):
    """ref:96-105"""
    rng = np.random.default_rng(101)
    for points in ((3, 3), (3, 5), (1, 5)):
        for shape in ((100, 100), (100, 75)):
            for order in (0, 1, 2, 3, 4):
                for mode in ALL_MODES:
                    if MODERN_SCIPY and mode in ("reflect", "nearest"):
                        continue
                    compare_with_scipy(ed, rng, shape, points, order=order, mode=mode)


def test_basic_3d(ed):
    """ref:107-111"""
    rng = np.random.default_rng(102)
    for points in ((3, 3, 3), (3, 5, 7), (1, 3, 5)):
        for shape in ((50, 50, 50), (100, 50, 25)):
            for order in (0, 1, 2, 3, 4):
                compare_with_scipy(ed, rng, shape, points, order=order)


def test_crop_2d(ed):
    """ref:113-120"""
    rng = np.random.default_rng(103)
    for lo, hi in ((0, 50), (20, 60), (50, 100)):
        for order in (0, 1, 2, 3, 4):
            compare_with_scipy(ed, rng, (100, 100), (3, 3), crop=(slice(lo, hi),) * 2, order=order)


def test_crop_3d(ed):
    """ref:122-127"""
    rng = np.random.default_rng(104)
    compare_with_scipy(ed, rng, (25, 25, 25), (3, 3, 5), crop=(slice(15, 25), slice(None), slice(None)), order=3)


def test_crop_rotate_zoom(ed):
    """ref:129-141: cropping commutes with the affine part (the output centre stays where it was)."""
    rng = np.random.default_rng(105)
    crop = (slice(10, 90), slice(20, 80))
    for rotate in (-30, 0, 30, None):
        for zoom in (0.5, 1.0, 1.5, None):
            for affine in (None, np.eye(3)):
                X = rng.random((100, 100))
                D = rng.standard_normal((2, 3, 3)) * 3
                whole = ed.deform_grid(X, D, rotate=rotate, zoom=zoom, affine=affine)
                part = ed.deform_grid(X, D, rotate=rotate, zoom=zoom, crop=crop, affine=affine)
                np.testing.assert_allclose(whole[crop], part, rtol=1e-05, atol=1e-08)


def _pair(v):
    return v if isinstance(v, list) else [v, v]


def test_multi_2d(ed):
    """ref:143-170: two inputs of different dtype with per-input order / mode / cval."""
    rng = np.random.default_rng(106)
    shape = (100, 75)
    for order in (0, 1, 2, 3, 4, [0, 3]):
        for crop in (None, (slice(15, 25), slice(15, 50))):
            for cval in (0.0, 1.0, [0.0, 1.0]):
                for mode in ("constant", ["constant", "reflect"]):
                    if MODERN_SCIPY and mode == ["constant", "reflect"]:
                        continue
                    D = rng.standard_normal((2, 3, 3)) * 25
                    A = rng.random(shape).astype("float64")
                    B = rng.random(shape).astype("float32")
                    o, m, c = _pair(order), _pair(mode), _pair(cval)
                    wantA = scipy_deform(A, D, order=o[0], crop=crop, cval=c[0], mode=m[0])
                    wantB = scipy_deform(B, D, order=o[1], crop=crop, cval=c[1], mode=m[1])
                    gotA, gotB = ed.deform_grid([A, B], D, order=order, crop=crop, cval=cval, mode=mode)
                    np.testing.assert_allclose(wantA, gotA, rtol=1e-05, atol=1e-06)
                    np.testing.assert_allclose(wantB, gotB, rtol=1e-05, atol=1e-06)


def test_multi_3d(ed):
    """ref:172-191"""
    rng = np.random.default_rng(107)
    shape = (25, 25, 30)
    for order in (0, 1, 2, 3, 4):
        for crop in (None, (slice(15, 20), slice(15, 25), slice(2, 10))):
            D = rng.standard_normal((3, 3, 3, 3)) * 25
            A, B = rng.random(shape), rng.random(shape)
            gotA, gotB = ed.deform_grid([A, B], D, order=order, crop=crop)
            np.testing.assert_allclose(scipy_deform(A, D, order=order, crop=crop), gotA, rtol=1e-05, atol=1e-08)
            np.testing.assert_allclose(scipy_deform(B, D, order=order, crop=crop), gotB, rtol=1e-05, atol=1e-08)


def test_different_strides(ed):
    """ref:193-208: same values, C and Fortran order, in one call (the reference only runs it; here the
    results are compared as well)."""
    rng = np.random.default_rng(108)
    A = rng.random((200, 150))
    B = np.array(A, order="F")
    assert A.strides != B.strides
    D = rng.standard_normal((2, 3, 3)) * 25
    gotA, gotB = ed.deform_grid([A, B], D, prefilter=False)
    want = scipy_deform(A, D, prefilter=False)
    np.testing.assert_allclose(want, gotA, rtol=1e-05, atol=1e-08)
    np.testing.assert_allclose(want, gotB, rtol=1e-05, atol=1e-08)


def test_axis(ed):
    """ref:210-251"""
    rng = np.random.default_rng(109)
    for shape, axis in (((30, 20, 3), (0, 1)), ((20, 3, 30), (0, 2)), ((100, 200, 3), (0, 1)),
                        ((200, 3, 100), (0, 2)), ((200, 3, 100, 4), (0, 2))):
        compare_with_scipy(ed, rng, shape, (3, 3), axis=axis)

    def both(A, B, D, axA, axB, crop=None):
        same = axA == axB
        gotA, gotB = ed.deform_grid([A, B], D, axis=axA if same else [axA, axB], crop=crop)
        np.testing.assert_allclose(scipy_deform(A, D, axis=axA, crop=crop), gotA, rtol=1e-05, atol=1e-08)
        np.testing.assert_allclose(scipy_deform(B, D, axis=axB, crop=crop), gotB, rtol=1e-05, atol=1e-08)

    # same axes in both inputs
    both(rng.random((3, 90, 80, 7)), rng.random((7, 90, 80)), rng.standard_normal((2, 5, 3)) * 25, (1, 2), (1, 2))
    # different axes
    both(rng.random((3, 20, 30)), rng.random((20, 30)), rng.standard_normal((2, 5, 3)) * 25, (1, 2), (0, 1))
    # with cropping
    A, B = rng.random((3, 90, 80, 7)), rng.random((7, 90, 80))
    D = rng.standard_normal((2, 5, 3)) * 25
    for crop in ((slice(30, 50), slice(20, 40)), (slice(0, 30), slice(0, 80))):
        both(A, B, D, (1, 2), (1, 2), crop=crop)


def test_grad_2d(ed):
    """ref:253-265: finite differences, every order x every mode."""
    rng = np.random.default_rng(110)
    for order in (0, 1, 2, 3, 4):
        for mode in ALL_MODES:
            X = rng.random((30, 25))
            D = rng.standard_normal((2, 3, 5)) * 3
            check_gradient_numerically(
                rng, X, lambda x: ed.deform_grid(x, D, order=order, mode=mode),
                lambda g, x: ed.deform_grid_gradient(g, D, order=order, mode=mode), n_tests=5)


def test_grad_crop(ed):
    """ref:267-279"""
    rng = np.random.default_rng(111)
    shape = (20, 20)
    for lo, hi in ((0, 10), (4, 12), (10, 20)):
        crop = (slice(lo, hi),) * 2
        X = rng.random(shape)
        D = rng.standard_normal((2, 3, 3)) * 3
        check_gradient_numerically(
            rng, X, lambda x: ed.deform_grid(x, D, crop=crop),
            lambda g, x: ed.deform_grid_gradient(g, D, crop=crop, X_shape=shape))


def test_grad_zoom(ed):
    """ref:281-293"""
    rng = np.random.default_rng(112)
    for zoom in (0.5, 1.0, 1.5):
        X = rng.random((30, 25))
        D = rng.standard_normal((2, 3, 5)) * 3
        check_gradient_numerically(
            rng, X, lambda x: ed.deform_grid(x, D, order=3, mode="constant", zoom=zoom),
            lambda g, x: ed.deform_grid_gradient(g, D, order=3, mode="constant", zoom=zoom), n_tests=5)


def test_grad_rotate(ed):
    """ref:295-307"""
    rng = np.random.default_rng(113)
    for rotate in (-20, 0, 20):
        X = rng.random((30, 25))
        D = rng.standard_normal((2, 3, 5)) * 3
        check_gradient_numerically(
            rng, X, lambda x: ed.deform_grid(x, D, order=3, mode="constant", rotate=rotate),
            lambda g, x: ed.deform_grid_gradient(g, D, order=3, mode="constant", rotate=rotate), n_tests=5)


def test_grad_with_list(ed):
    """ref:309-338: a list call of the gradient equals the item-by-item calls."""
    rng = np.random.default_rng(114)
    shape = (100, 75)
    for order in (0, 1, 2, 3, 4, [0, 3]):
        for crop in (None, (slice(15, 25), slice(15, 50))):
            for cval in (0.0, 1.0, [0.0, 1.0]):
                for mode in ("constant", ["constant", "reflect"]):
                    D = rng.standard_normal((2, 3, 3)) * 25
                    A = rng.random(shape).astype("float64")
                    B = rng.random(shape).astype("float32")
                    YA, YB = ed.deform_grid([A, B], D, order=order, crop=crop, cval=cval, mode=mode)
                    gA = rng.random(YA.shape).astype("float64")
                    gB = rng.random(YB.shape).astype("float32")
                    o, m, c = _pair(order), _pair(mode), _pair(cval)
                    oneA = ed.deform_grid_gradient(gA, D, order=o[0], crop=crop, cval=c[0], mode=m[0], X_shape=A.shape)
                    oneB = ed.deform_grid_gradient(gB, D, order=o[1], crop=crop, cval=c[1], mode=m[1], X_shape=B.shape)
                    bothA, bothB = ed.deform_grid_gradient([gA, gB], D, order=order, crop=crop, cval=cval, mode=mode,
                                                           X_shape=[A.shape, B.shape])
                    np.testing.assert_allclose(oneA, bothA, rtol=1e-05, atol=1e-08)
                    # float32: the device sums the contributions to a cell with float atomics, in an order that differs
                    # from launch to launch, so two calls agree to float32 accumulation noise, not bit for bit as the
                    # single-threaded reference does: same rtol, atol scaled to the largest entry (the prefilter adjoint
                    # leaves entries of ~1e-7 next to entries of ~1 by cancellation)
                    np.testing.assert_allclose(oneB, bothB, rtol=1e-05, atol=2e-06 * max(1.0, float(np.abs(oneB).max())))


def _torch_case(ed, rng, shape, points, device, order=3, sigma=25, crop=None, mode="constant"):
    """ref:470-500 (single tensor) -- on ``device``; the NumPy API of the same package is the expected value."""
    import torch
    import elasticdeform.torch as etorch
    D = rng.standard_normal((len(shape),) + tuple(points)) * sigma
    Xv = rng.random(shape)
    Yref = ed.deform_grid(Xv, D, order=order, crop=crop, mode=mode)
    gv = rng.random(Yref.shape)
    dXref = ed.deform_grid_gradient(gv, D, order=order, crop=crop, mode=mode, X_shape=shape)
    X = torch.tensor(Xv, requires_grad=True, device=device)
    Y = etorch.deform_grid(X, torch.tensor(D), order=order, crop=crop, mode=mode)
    assert Y.device.type == device
    Y.backward(torch.tensor(gv, device=device))
    np.testing.assert_almost_equal(Yref, Y.detach().cpu().numpy())
    np.testing.assert_almost_equal(dXref, X.grad.detach().cpu().numpy())


@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_basic_2d_torch(ed, device):
    """ref:462-468 (CPU tensors as in the reference) and the same cases with CUDA tensors."""
    rng = np.random.default_rng(115)
    for order in (0, 1, 2):
        for crop in (None, (slice(20, 80), slice(30, 70))):
            for mode in ALL_MODES:
                _torch_case(ed, rng, (100, 100), (3, 3), device, order=order, mode=mode, crop=crop)


@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_multi_2d_torch(ed, device):
    """ref:502-568: two tensors through one call, tuple in -> tuple out."""
    import torch
    import elasticdeform.torch as etorch
    rng = np.random.default_rng(116)
    shape = (100, 75)
    for order in (0, 1, 2, 3, 4, [0, 3]):
        for crop in (None, (slice(15, 25), slice(15, 50))):
            for mode in ("constant", ["constant", "reflect"]):
                D = rng.standard_normal((2, 3, 3)) * 25
                Av, Bv = rng.random(shape), rng.random(shape)
                YA, YB = ed.deform_grid([Av, Bv], D, order=order, crop=crop, mode=mode)
                gA, gB = rng.random(YA.shape), rng.random(YB.shape)
                dA, dB = ed.deform_grid_gradient([gA, gB], D, order=order, crop=crop, mode=mode, X_shape=[shape, shape])
                A = torch.tensor(Av, requires_grad=True, device=device)
                B = torch.tensor(Bv, requires_grad=True, device=device)
                TA, TB = etorch.deform_grid([A, B], torch.tensor(D), order=order, crop=crop, mode=mode)
                TA.backward(torch.tensor(gA, device=device), retain_graph=True)
                TB.backward(torch.tensor(gB, device=device))
                np.testing.assert_almost_equal(YA, TA.detach().cpu().numpy())
                np.testing.assert_almost_equal(YB, TB.detach().cpu().numpy())
                np.testing.assert_almost_equal(dA, A.grad.detach().cpu().numpy())
                np.testing.assert_almost_equal(dB, B.grad.detach().cpu().numpy())
