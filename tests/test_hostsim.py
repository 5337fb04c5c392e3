"""The per-voxel device code, compiled for the host (tests/hostsim.py), against the
oracle -- CPU only.  This is what lets kernel logic be debugged without a GPU; the
`-m gpu` suite then only has to confirm that nvcc's build of the same code agrees.
"""
import numpy as np
import pytest

from common import golden_names, load_golden, call_args, as_list, grad_kwargs
from oracle import oracle as O
import hostsim


@pytest.fixture(scope="module", autouse=True)
def _register():
    hs = hostsim.HostSimModule()
    O.register_backend("hostsim", hs, hs.spline_filter1d)


@pytest.mark.parametrize("name", golden_names())
def test_generic_device_code_matches_golden_bit_exact(name):
    g = load_golden(name)
    Y = as_list(O.deform_grid(call_args(g), g["D"], impl="hostsim", **g["kwargs"]))
    for y, yref in zip(Y, g["Y"]):
        np.testing.assert_array_equal(y, yref)
    dX = as_list(O.deform_grid_gradient(call_args(g, "dY"), g["D"], impl="hostsim", **grad_kwargs(g)))
    for dx, dxref in zip(dX, g["dX"]):
        np.testing.assert_array_equal(dx, dxref)      # sequential host adds: same order as the reference


def _coords_case(rng, shape, pts, order, mode, sigma, crop_off=None, affine=None, oshape=None, zero=False):
    X = rng.random(shape).astype("float32")
    D = np.zeros((len(shape),) + pts) if zero else rng.standard_normal((len(shape),) + pts) * sigma
    Df = O._prefilter_displacement(D, O._backend("port")[1])
    out = np.zeros(oshape or shape, dtype="float32")
    args = ([X], Df, crop_off, [out], [tuple(range(len(shape)))], [order], [O.MODES[mode]], [0.0], affine)
    f = hostsim.fast_coords(*args)
    e = hostsim.fast_coords(*args, exact=True)
    return f, e, out.size


@pytest.mark.parametrize("mode", list(O.MODES))
@pytest.mark.parametrize("order", [0, 1, 2, 3])
def test_fast_coordinate_pipeline_discrete_decisions_are_exact(mode, order):
    """Separable displacement + near-threshold re-evaluation must give the SAME window starts and
    constant flags as the reference-order evaluation of every voxel; fractions within 1e-6."""
    rng = np.random.default_rng(100 + order)
    f, e, n = _coords_case(rng, (24, 30, 70), (5, 4, 5), order, mode, 5.0)
    np.testing.assert_array_equal(f[0], e[0])
    np.testing.assert_array_equal(f[2], e[2])
    assert np.abs(f[1] - e[1]).max() < 1e-6
    assert f[3] <= max(10, n // 10000)                # the exact path is rare


def test_fast_coordinates_2d_reflect_cfg1_shape():
    rng = np.random.default_rng(5)
    f, e, n = _coords_case(rng, (200, 300), (3, 3), 3, "reflect", 25.0)
    np.testing.assert_array_equal(f[0], e[0])
    np.testing.assert_array_equal(f[2], e[2])


def test_fast_coordinates_on_the_integer_lattice():
    """Zero displacement / half-integer translations put EVERY voxel on a rounding threshold."""
    rng = np.random.default_rng(6)
    f, e, n = _coords_case(rng, (20, 24, 30), (3, 3, 3), 0, "wrap", 0.0, zero=True,
                           affine=np.array([[1., 0, 0, 0.5], [0, 1, 0, 1.0], [0, 0, 1, -0.5]]))
    np.testing.assert_array_equal(f[0], e[0])
    np.testing.assert_array_equal(f[2], e[2])
    assert f[3] == 0                                   # all-zero control points: no re-evaluation needed
    # constant displacement of exactly 0.5: every voxel needs (and gets) the exact order
    X = rng.random((16, 18, 20)).astype("float32")
    D = np.full((3, 3, 3, 3), 0.5)
    Df = O._prefilter_displacement(D, O._backend("port")[1])
    out = np.zeros_like(X)
    args = ([X], Df, None, [out], [(0, 1, 2)], [0], [O.MODES["nearest"]], [0.0], None)
    f = hostsim.fast_coords(*args)
    e = hostsim.fast_coords(*args, exact=True)
    np.testing.assert_array_equal(f[0], e[0])
    assert f[3] > 0


def test_fast_coordinates_crop_affine():
    rng = np.random.default_rng(7)
    A = np.array([[1.0, 0, 0, 0.], [0, 1.1, 0.1, -2.], [0, -0.1, 1.1, 3.]])
    f, e, n = _coords_case(rng, (40, 40, 40), (5, 5, 5), 3, "constant", 2.0,
                           crop_off=np.array([10, 10, 10]), affine=A, oshape=(20, 20, 20))
    np.testing.assert_array_equal(f[0], e[0])
    np.testing.assert_array_equal(f[2], e[2])
    assert np.abs(f[1] - e[1]).max() < 1e-6


def test_line_filters_device_code():
    import scipy.ndimage
    rng = np.random.default_rng(8)
    hs = hostsim.HostSimModule()
    for order in (2, 3, 4, 5):
        for n in (1, 2, 5, 30, 129):
            x = rng.standard_normal((n, 3)).astype("float32")
            a = np.zeros_like(x)
            scipy.ndimage.spline_filter1d(x, axis=0, order=order, output=a)
            b = np.zeros_like(x)
            hs.spline_filter1d(x, 0, order, b)
            np.testing.assert_array_equal(a, b)
            c = np.zeros_like(x)
            hs.spline_filter1d_grad(x, c, 0, order)
            np.testing.assert_array_equal(c, O.spline_filter1d_grad(x, 0, order, impl="port"))


def test_line_filters_bit_exact_all_lengths():
    """The register-blocked line recursions of csrc/edf_spline_lines.h (blocks of 8 elements, remainders, lines shorter
    than a block): K3 == scipy.ndimage.spline_filter1d and K4 == the reference's spline_filter1d_grad bit for bit."""
    import scipy.ndimage
    H = hostsim.HostSimModule()
    mod, _ = O._backend("ref" if O.ref_available() else "port")
    rng = np.random.default_rng(0)
    for n in (2, 3, 5, 8, 9, 10, 17, 18, 25, 31, 64, 100, 257):
        for order in (2, 3, 4, 5):
            for dt in (np.float64, np.float32):
                X = rng.random((3, n, 4)).astype(dt)
                out = np.zeros_like(X)
                H.spline_filter1d(X, 1, order, out)
                ref = scipy.ndimage.spline_filter1d(X, order=order, axis=1, output=dt, mode="mirror")
                np.testing.assert_array_equal(out, ref, err_msg="K3 n=%d order=%d" % (n, order))
                out2 = np.zeros_like(X)
                H.spline_filter1d_grad(X, out2, 1, order)
                ref2 = np.zeros_like(X)
                mod.spline_filter1d_grad(X, ref2, 1, order)
                np.testing.assert_array_equal(out2, ref2, err_msg="K4 n=%d order=%d" % (n, order))
