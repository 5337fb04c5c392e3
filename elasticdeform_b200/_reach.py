"""Rigorous per-slab bounds of the displacement along the first deformed axis (host side, NumPy only).

The slab pipeline of ``deform_grid.py`` overlaps upload, kernel and download of different slabs of a
host volume.  It may launch an output slab as soon as every input plane that slab can read has
arrived, so it needs a bound -- a proof, not an estimate -- of the displacement d_0 over the slab.

d_0 is a tensor-product cubic B-spline of the prefiltered control coefficients (deform.c:650-758):
the value at control position cp is sum_i c[m(i)] * beta3(cp - i) over the four taps around cp, with
tap indices outside [0, P) folded back by the whole-sample mirror map m (deform.c:664-686).  B-spline
weights are non-negative and sum to one, so d_0 lies in the convex hull of the coefficients it
touches.  For the raw prefiltered coefficients that hull is useless: the prefilter overshoots (a
random 5^3 grid of sigma 8 gives coefficients up to ~230 for a field that never exceeds ~30).  Knot
insertion fixes that.  The two-scale relation of the uniform cubic B-spline,

    c'[2i] = (c[i-1] + 6 c[i] + c[i+1]) / 8,        c'[2i+1] = (c[i] + c[i+1]) / 2,

rewrites the SAME function on a grid of half the spacing; its coefficients again bound the function
and their distance to it shrinks 4x per level.  After r levels a slab only touches the refined
coefficients whose support overlaps the slab's range of control positions, which gives a separate
(lower, upper) bound per slab.  Everything is linear in c, so mirror extension + r refinements are
one cached matrix per (P, r) and the whole bound is naxis small matrix products.
"""
import numpy

_MATRIX_CACHE = {}
_MATRIX_LOCK = __import__('threading').Lock()
_MAX_FINE = 1 << 18                 # cap of the refined coefficient count (levels are lowered to fit)


def _refine_rows(v):
    """One knot-insertion step along axis 0 of v (first row = index -1 before and after)."""
    even = (v[:-2] + 6.0 * v[1:-1] + v[2:]) * 0.125
    odd = (v[:-1] + v[1:]) * 0.5
    out = numpy.empty((even.shape[0] + odd.shape[0],) + v.shape[1:], dtype=v.dtype)
    out[0::2] = odd
    out[1::2] = even
    return out


def refine_matrix(P, r):
    """(n_fine, P) matrix taking P control coefficients to the level-r refined coefficients with
    indices -1 .. 2^r * P + 1 (units of 2^-r control spacings), mirror extension included."""
    key = (int(P), int(r))
    with _MATRIX_LOCK:
        M = _MATRIX_CACHE.get(key)
    if M is None:
        idx = numpy.arange(-1, P + 2)
        if P > 1:
            s2 = 2 * (P - 1)
            m = numpy.abs(idx) % s2
            m = numpy.where(m >= P, s2 - m, m)
        else:
            m = numpy.zeros_like(idx)
        M = numpy.eye(P)[m]
        for _ in range(r):
            M = _refine_rows(M)
        with _MATRIX_LOCK:
            if len(_MATRIX_CACHE) > 64:
                _MATRIX_CACHE.clear()
            _MATRIX_CACHE[key] = M
    return M


def choose_level(points, want=3):
    """Highest refinement level <= want whose refined grid stays below _MAX_FINE values."""
    r = want
    while r > 0:
        n = 1
        for P in points:
            n *= (P << r) + 3
        if n <= _MAX_FINE:
            break
        r -= 1
    return r


def plane_hull(coef0, level=None):
    """Per refined index along axis 0: (min, max) over all other axes of the refined coefficients of
    the axis-0 displacement component.  Returns (lo[n0], hi[n0], level)."""
    c = numpy.asarray(coef0, dtype=numpy.float64)
    r = choose_level(c.shape, 4 if max(c.shape) <= 3 else 3) if level is None else int(level)
    with numpy.errstate(invalid='ignore', over='ignore'):      # non-finite coefficients: the caller checks the result
        if c.ndim == 3:
            # the common case as three plain matrix products (half the time of the generic contraction)
            M0, M1, M2 = (refine_matrix(P, r) for P in c.shape)
            F = (M0 @ c.reshape(c.shape[0], -1)).reshape(M0.shape[0], c.shape[1], c.shape[2])
            F = numpy.matmul(M1, F @ M2.T)
        else:
            F = c
            for a in range(c.ndim):
                F = numpy.moveaxis(numpy.tensordot(refine_matrix(c.shape[a], r), F, axes=([1], [a])), 0, a)
        F = F.reshape(F.shape[0], -1)
        return F.min(axis=1), F.max(axis=1), r


def slab_bounds(coef0, dim0, offset0, slabs, level=None):
    """Bounds of d_0 over output slabs.

    coef0    prefiltered control coefficients of the axis-0 displacement component, shape (P_0, ...)
    dim0     extent of the deformed (input) volume along axis 0 (the I_0 of deform.c:655)
    offset0  crop offset along axis 0
    slabs    list of (a, b): output planes [a, b) of each slab

    Returns a list of (dmin, dmax) floats per slab with dmin <= d_0 <= dmax for every voxel of the
    slab, or None when the coefficients are not finite.
    """
    lo, hi, r = plane_hull(coef0, level)
    if not (numpy.all(numpy.isfinite(lo)) and numpy.all(numpy.isfinite(hi))):
        return None
    P0 = int(numpy.shape(coef0)[0])
    n0 = lo.shape[0]
    lo, hi = lo.tolist(), hi.tolist()
    scale = float(1 << r) * (P0 - 1) / float(dim0 - 1) if dim0 > 1 else 0.0
    out = []
    for (a, b) in slabs:
        sa = scale * (a + offset0)
        sb = scale * (b - 1 + offset0)
        # basis j is non-zero on (j - 2, j + 2): the taps of s are floor(s) - 1 .. floor(s) + 2; one more
        # index on either side absorbs the rounding of s itself.  Array position = index + 1.
        ja = max(0, int(sa // 1) - 1)
        jb = min(n0, int(sb // 1) + 5)
        if ja >= jb:                                  # cannot happen for positions inside [0, P_0 - 1]
            ja, jb = 0, n0
        out.append((min(lo[ja:jb]), max(hi[ja:jb])))
    return out


def integer_reach(bounds, order):
    """(lo, hi) per slab: every input plane a voxel of the slab reads or scatters to, taps included, lies in
    [o + offset0 + lo, o + offset0 + hi] (o = the voxel's output plane).  The window of an order-n spline
    starts at floor(c) - n/2 or floor(c + 0.5) - n/2 and spans n + 1 taps (deform.c:784-788): n + 2 planes on
    either side of the displaced coordinate cover it with room to spare."""
    pad = int(order) + 2
    return [(int(numpy.floor(lo)) - pad, int(numpy.ceil(hi)) + pad) for lo, hi in bounds]


def forward_waits(slabs, reach, in0, offset0, h):
    """Forward pipeline: index of the last UPLOAD slab (height h, in0 input planes) that output slab k has to
    wait for -- the slab holding the highest input plane it can read."""
    return [min(in0 - 1, max(0, b - 1 + offset0 + hi)) // h for (a, b), (_, hi) in zip(slabs, reach)]


def gradient_final_slabs(slabs, reach, in0, offset0, h):
    """Gradient pipeline: after output slab k has been scattered, how many leading dX slabs (height h) are
    final, i.e. below the lowest input plane any LATER output slab can still add to.  Non-decreasing in k;
    all of them after the last slab."""
    n_out = len(slabs)
    n_in = -(-in0 // h)
    lowest = [in0] * (n_out + 1)
    for k in range(n_out - 1, -1, -1):
        lowest[k] = min(lowest[k + 1], slabs[k][0] + offset0 + reach[k][0])
    out, done = [], 0
    for k in range(n_out):
        j = n_in if k == n_out - 1 else min(n_in, max(0, lowest[k + 1]) // h)
        done = max(done, j)
        out.append(done)
    return out
