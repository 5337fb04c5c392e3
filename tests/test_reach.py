"""The slab pipeline's displacement bound (elasticdeform_b200/_reach.py) must be a proof: for every slab
the true displacement along the first axis -- evaluated here voxel by voxel with SciPy's cubic spline
in 'mirror' mode, the restatement the reference's own tests use (tests/test_deform_grid.py:36-72) --
lies inside the bound, and the bound is tight enough to be useful."""
import numpy as np
import pytest
import scipy.ndimage as ndi

from elasticdeform_b200 import _reach

CASES = [((5, 5, 5), (64, 48, 40), 8.0), ((3, 3, 3), (64, 32, 32), 25.0), ((5, 4, 7), (96, 20, 33), 8.0),
         ((1, 3, 2), (40, 10, 10), 5.0), ((2, 2, 2), (33, 9, 9), 10.0), ((8, 8), (128, 64), 12.0),
         ((3, 4, 2, 3), (32, 8, 8, 8), 6.0), ((4,), (100,), 20.0), ((12, 3, 3), (80, 16, 16), 15.0)]


def _field(P, N, sigma, seed):
    rng = np.random.default_rng(seed)
    c = rng.standard_normal(P) * sigma
    for a in range(len(P)):
        if P[a] > 1:
            c = ndi.spline_filter1d(c, 3, axis=a, mode='mirror')
    grids = np.meshgrid(*[np.arange(n) * ((p - 1) / (n - 1)) for n, p in zip(N, P)], indexing='ij')
    return c, ndi.map_coordinates(c, grids, order=3, mode='mirror', prefilter=False)


@pytest.mark.parametrize("P,N,sigma", CASES)
@pytest.mark.parametrize("h", [1, 8, 19])
def test_slab_bounds_contain_the_field(P, N, sigma, h):
    for seed in range(3):
        c, d = _field(P, N, sigma, seed)
        slabs = [(a, min(N[0], a + h)) for a in range(0, N[0], h)]
        bounds = _reach.slab_bounds(c, N[0], 0, slabs)
        assert len(bounds) == len(slabs)
        for (a, b), (lo, hi) in zip(slabs, bounds):
            assert lo <= d[a:b].min() + 1e-9 and hi >= d[a:b].max() - 1e-9, (P, seed, a, b)


def test_bounds_with_crop_offset_and_tightness():
    P, N = (5, 5, 5), (128, 40, 40)
    c, d = _field(P, N, 8.0, 7)
    off, out0, h = 20, 90, 16
    slabs = [(a, min(out0, a + h)) for a in range(0, out0, h)]
    bounds = _reach.slab_bounds(c, N[0], off, slabs)
    slack = 0.0
    for (a, b), (lo, hi) in zip(slabs, bounds):
        part = d[a + off:b + off]
        assert lo <= part.min() + 1e-9 and hi >= part.max() - 1e-9
        slack = max(slack, part.min() - lo, hi - part.max())
    # the hull of the raw prefiltered coefficients overshoots by far more than the field's own range;
    # the refined hull stays within a fraction of it
    assert np.abs(c).max() > 2 * np.abs(d).max()
    assert slack < 0.5 * (d.max() - d.min())


def test_refinement_reproduces_the_same_spline():
    # level-r coefficients describe the same function: evaluate both at random positions
    rng = np.random.default_rng(3)
    P = 6
    c = rng.standard_normal(P)
    t = rng.random(50) * (P - 1)
    ref = ndi.map_coordinates(c, [t], order=3, mode='mirror', prefilter=False)
    for r in (0, 1, 3):
        f = _reach.refine_matrix(P, r) @ c                     # indices -1 .. 2^r P + 1
        s = t * (1 << r)
        val = np.zeros_like(t)
        for j in range(f.shape[0]):
            x = np.abs(s - (j - 1))
            w = np.where(x < 1, 2 / 3 - x * x + x ** 3 / 2, np.where(x < 2, (2 - x) ** 3 / 6, 0.0))
            val += f[j] * w
        np.testing.assert_allclose(val, ref, atol=1e-12)


def test_non_finite_coefficients_are_refused():
    c = np.zeros((3, 3, 3))
    c[1, 1, 1] = np.inf
    assert _reach.slab_bounds(c, 64, 0, [(0, 32), (32, 64)]) is None
    c[1, 1, 1] = np.nan
    assert _reach.slab_bounds(c, 64, 0, [(0, 32), (32, 64)]) is None


@pytest.mark.parametrize("sigma,order,off,out0,h", [(8.0, 3, 0, 64, 8), (25.0, 3, 0, 64, 8), (8.0, 0, 10, 40, 7),
                                                   (30.0, 1, 5, 50, 8), (3.0, 5, 0, 64, 16)])
def test_pipeline_schedules_respect_the_true_field(sigma, order, off, out0, h):
    """forward_waits / gradient_final_slabs against the voxel-by-voxel field: an output slab never reads an
    input plane that has not been uploaded when it starts, and a dX slab is never sent home while a later
    output slab can still add to it (plane ranges include the spline's taps; coordinates outside the volume
    are clamped to it, as 'nearest' does -- 'constant' touches nothing there)."""
    P, N = (5, 4, 5), (64, 24, 20)
    c, d = _field(P, N, sigma, 11)
    in0 = N[0]
    slabs = [(a, min(out0, a + h)) for a in range(0, out0, h)]
    bounds = _reach.slab_bounds(c, in0, off, slabs)
    reach = _reach.integer_reach(bounds, order)
    o = np.arange(out0)[:, None, None] + off
    src = o + d[off:off + out0]
    first = np.floor(src).astype(int) - order // 2 - 1              # generous: one plane beyond the window start
    last = first + order + 2
    first, last = np.clip(first, 0, in0 - 1), np.clip(last, 0, in0 - 1)
    waits = _reach.forward_waits(slabs, reach, in0, off, h)
    final = _reach.gradient_final_slabs(slabs, reach, in0, off, h)
    n_in = -(-in0 // h)
    assert final[-1] == n_in and all(final[k] <= final[k + 1] for k in range(len(final) - 1))
    for k, (a, b) in enumerate(slabs):
        assert 0 <= waits[k] < n_in
        assert last[a:b].max() < (waits[k] + 1) * h, (k, "forward slab starts before its inputs arrived")
        if k + 1 < len(slabs):
            assert first[b:].min() >= final[k] * h, (k, "dX slab downloaded before its last contribution")
    # and the schedules are not vacuous: for the gentle field something overlaps
    if sigma <= 8.0 and out0 == 64 and h == 8:
        assert waits[0] < n_in - 1 and final[len(slabs) // 2] > 0
