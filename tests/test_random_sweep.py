"""Seeded random sweep (CPU): the host-compiled device code against the oracle port over random
shapes / grids / orders / modes / dtypes / crops / affines / axis selections, forward and gradient.
The reference's own tests draw unseeded random cases of exactly this kind (tests/test_deform_grid.py:
run_comparison, test_multi_2d, test_axis); here they are seeded and checked bit for bit."""
import numpy as np
import pytest

from oracle import oracle as O
import hostsim


@pytest.fixture(scope="module", autouse=True)
def _register():
    hs = hostsim.HostSimModule()
    O.register_backend("hostsim", hs, hs.spline_filter1d)


MODES = list(O.MODES)
DTYPES = ["float64", "float32", "int32", "uint8", "int16", "uint16", "int64"]


def _case(rng):
    naxis = int(rng.integers(1, 4))
    shape = tuple(int(rng.integers(4, 22)) for _ in range(naxis))
    points = tuple(int(rng.integers(1, 6)) for _ in range(naxis))
    extra_front = int(rng.integers(0, 2))
    extra_back = int(rng.integers(0, 2))
    full = (3,) * extra_front + shape + (2,) * extra_back
    axis = tuple(range(extra_front, extra_front + naxis)) if (extra_front or extra_back) else None
    kw = dict(order=int(rng.integers(0, 6)), mode=MODES[int(rng.integers(0, 5))], cval=float(rng.normal() * 3))
    if axis is not None:
        kw["axis"] = axis
    if rng.random() < 0.4:
        crop = []
        for n in shape:
            a = int(rng.integers(0, max(1, n // 2)))
            b = int(rng.integers(a + 1, n + 1))
            crop.append(slice(a, b))
        kw["crop"] = tuple(crop)
    if rng.random() < 0.4:
        A = np.eye(naxis) + rng.normal(size=(naxis, naxis)) * 0.1
        kw["affine"] = np.concatenate([A, rng.normal(size=(naxis, 1)) * 2], axis=1)
    if rng.random() < 0.3:
        kw["prefilter"] = False
    dt = DTYPES[int(rng.integers(0, len(DTYPES)))]
    X = (rng.random(full) * 120 - 10).astype(dt)
    D = rng.standard_normal((naxis,) + points) * float(rng.choice([0.0, 1.0, 4.0, 12.0]))
    return X, D, kw


@pytest.mark.parametrize("seed", range(40))
def test_random_case(seed):
    rng = np.random.default_rng(7000 + seed)
    for _ in range(3):
        X, D, kw = _case(rng)
        a = O.deform_grid(X, D, impl="port", **kw)
        b = O.deform_grid(X, D, impl="hostsim", **kw)
        np.testing.assert_array_equal(a, b, err_msg=repr(kw))
        G = (rng.random(a.shape) * 7).astype(X.dtype)
        gkw = dict(kw)
        gkw["X_shape"] = X.shape
        ga = O.deform_grid_gradient(G, D, impl="port", **gkw)
        gb = O.deform_grid_gradient(G, D, impl="hostsim", **gkw)
        np.testing.assert_array_equal(ga, gb, err_msg=repr(kw))


@pytest.mark.parametrize("seed", range(12))
def test_random_fast_coordinates(seed):
    """Fast coordinate pipeline == reference-order coordinates in every discrete decision."""
    rng = np.random.default_rng(9000 + seed)
    naxis = int(rng.integers(2, 4))
    shape = tuple(int(rng.integers(20, 70)) for _ in range(naxis))
    points = tuple(int(rng.integers(2, 5)) for _ in range(naxis))
    order = int(rng.integers(0, 6))
    mode = MODES[int(rng.integers(0, 5))]
    X = rng.random(shape).astype("float32")
    D = rng.standard_normal((naxis,) + points) * float(rng.choice([0.5, 4.0, 10.0]))
    Df = O._prefilter_displacement(D, O._backend("port")[1])
    out = np.zeros(shape, dtype="float32")
    aff = None
    if rng.random() < 0.5:
        A = np.eye(naxis) + rng.normal(size=(naxis, naxis)) * 0.05
        aff = np.concatenate([A, rng.normal(size=(naxis, 1))], axis=1)
    args = ([X], Df, None, [out], [tuple(range(naxis))], [order], [O.MODES[mode]], [0.0], aff)
    try:
        f = hostsim.fast_coords(*args)
    except RuntimeError:
        pytest.skip("control grid too dense for the fast tables")
    e = hostsim.fast_coords(*args, exact=True)
    np.testing.assert_array_equal(f[0], e[0])
    np.testing.assert_array_equal(f[2], e[2])
    assert np.abs(f[1] - e[1]).max() < 1e-6
