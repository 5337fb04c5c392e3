"""Batched entry: many independent volumes, each with its own displacement / affine.

The reference has no batch dimension -- a batch is a Python loop around
``deform_grid`` (reference README.md:117-133), one single-threaded C call per
volume.  Here the whole batch is handed to the C-ABI in ONE call
(``edf_deform_grid_batch``), which enqueues the kernels back to back on one stream,
and batches shard across GPUs by volume (one process per GPU, no exchange).
"""
import ctypes
import importlib

import numpy

from . import _lib

_dg = importlib.import_module(__package__ + ".deform_grid")


def shard_range(n_items, rank, world_size):
    """Contiguous block of `n_items` independent volumes owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(int(n_items), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def deform_grid_batch(Xs, displacements, order=3, mode='constant', cval=0.0, crop=None,
                      prefilter=True, affines=None, gradient=False, X_shape=None, _flags=0):
    """Deform every volume Xs[b] with its own displacements[b] (and affines[b]).

    Same per-volume semantics as ``deform_grid`` (or ``deform_grid_gradient`` with
    ``gradient=True``, where Xs are the upstream gradients and ``X_shape`` the common
    input shape).  Volumes may be NumPy arrays or CUDA tensors; returns a list.
    """
    torch = _dg.torch
    lib = _dg._require_cuda()
    nb = len(Xs)
    assert len(displacements) == nb, 'one displacement per volume'
    if affines is not None:
        assert len(affines) == nb, 'one affine per volume'
    if nb == 0:
        return []
    device = _dg._device_of(Xs)
    problems = (_lib.EdfProblem * nb)()
    keep, outs, dxs = [], [], []
    # everything that does not depend on the volume is normalised once when the batch is uniform (same shape and
    # dtype throughout: the data-augmentation case); the per-volume Python work is then the inverse affine (when the
    # affines differ) and one edf_problem
    uniform = all(tuple(x.shape) == tuple(Xs[0].shape) and x.dtype == Xs[0].dtype for x in Xs)
    shared = None
    with torch.cuda.device(device):
        d_all = _prefilter_displacements_stacked(lib, displacements, Xs[0], device)
        if uniform and d_all is not None and not (prefilter and int(_dg._normalize_order(order, [0])[0]) > 1):
            res = _uniform_batch(lib, device, Xs, d_all, order, mode, cval, crop, affines, gradient, X_shape, _flags)
            if res is not None:
                return res
        inv_cache = {}
        for b in range(nb):
            X = [Xs[b]]
            if shared is None or not uniform:
                if gradient:
                    shp = tuple(X_shape) if X_shape is not None else tuple(Xs[b].shape)
                    meta = [_dg._ShapeOnly(shp)]
                else:
                    shp = None
                    meta = X
                axis, deform_shape = _dg._normalize_axis_list(None, meta)
                out_shapes, offset = _dg._compute_output_shapes(meta, axis, deform_shape, crop)
                od = _dg._normalize_order(order, X)
                md = _dg._normalize_mode(mode, X)
                cv = _dg._normalize_cval(cval, X)
                shared = (shp, axis, out_shapes, offset, od, md, cv)
            shp, axis, out_shapes, offset, od, md, cv = shared
            aff = None if affines is None else affines[b]
            key = id(aff)
            if key not in inv_cache:
                inv_cache[key] = _dg._compute_inverse_affine(_dg._normalize_affine(aff, axis))
            inv = inv_cache[key]
            if d_all is not None:
                d_f = d_all[b]
            else:
                d_f = _dg._prefilter_displacement(lib, _dg._normalize_displacement(displacements[b], X, axis), device)
            xd = _dg._to_device(Xs[b], device)
            if gradient:
                if tuple(out_shapes[0]) != tuple(xd.shape):
                    raise ValueError("X_shape does not match output shape and cropping.")
                dx = torch.zeros(shp, dtype=xd.dtype, device=device)
                pr, k = _dg._build_problem([dx], [xd], d_f, offset, axis, od, md, cv, inv, _flags)
                dxs.append(dx)
            else:
                src = xd
                if prefilter and od[0] > 1:
                    x_f = torch.empty_like(xd)
                    for ax in axis[0]:
                        _dg._spline_filter1d_device(lib, src, x_f, ax, int(od[0]))
                        src = x_f
                out = torch.empty(tuple(out_shapes[0]), dtype=xd.dtype, device=device)
                pr, k = _dg._build_problem([src], [out], d_f, offset, axis, od, md, cv, inv, _flags)
                outs.append(out)
            problems[b] = pr
            keep.append(k)
        _lib.check(lib.edf_deform_grid_batch(problems, nb, 1 if gradient else 0, _dg._stream_ptr(device)))
        if gradient:
            res = []
            for b, dx in enumerate(dxs):
                if prefilter and int(_dg._normalize_order(order, [0])[0]) > 1:
                    x_f = torch.empty_like(dx)
                    src = dx
                    for ax in range(dx.ndim):
                        _dg._spline_filter1d_device(lib, src, x_f, ax, int(_dg._normalize_order(order, [0])[0]), adjoint=True)
                        src = x_f
                    dx = x_f
                res.append(_dg._from_device(dx, Xs[b]))
            return res
        return [_dg._from_device(o, x) for o, x in zip(outs, Xs)]


def inverse_affines_stacked(affines, axis):
    """(nb, n, n + 1) inverse (output -> input) maps of a batch of forward affine maps in one vectorised inversion;
    row b equals ``_compute_inverse_affine(_normalize_affine(affines[b], axis))`` (reference deform_grid.py:382-399)."""
    naxis = len(axis[0])
    A = numpy.stack([_dg._normalize_affine(a, axis) for a in affines])              # (nb, n, n + 1)
    Ainv = numpy.linalg.inv(A[:, :, :naxis])
    inv_all = numpy.concatenate([Ainv, -numpy.einsum('bij,bj->bi', Ainv, A[:, :, naxis])[:, :, None]], axis=2)
    return numpy.ascontiguousarray(inv_all, dtype='float64')


def _uniform_batch(lib, device, Xs, d_all, order, mode, cval, crop, affines, gradient, X_shape, flags):
    """Volumes of one shape and dtype without a per-volume prefilter: ONE edf_problem describes the batch and the
    C-ABI receives three arrays of device addresses (edf_deform_grid_batch_uniform); the per-volume Python work is
    reading a data pointer.  Returns None when the batch does not qualify (strided views with different layouts)."""
    torch = _dg.torch
    nb = len(Xs)
    X0 = [Xs[0]]
    if gradient:
        shp = tuple(X_shape) if X_shape is not None else tuple(Xs[0].shape)
        meta = [_dg._ShapeOnly(shp)]
    else:
        shp = None
        meta = X0
    axis, deform_shape = _dg._normalize_axis_list(None, meta)
    out_shapes, offset = _dg._compute_output_shapes(meta, axis, deform_shape, crop)
    od = _dg._normalize_order(order, X0)
    md = _dg._normalize_mode(mode, X0)
    cv = _dg._normalize_cval(cval, X0)
    naxis = len(axis[0])
    # inverse affine maps of the whole batch in one vectorised inversion
    inv_all = None if affines is None else inverse_affines_stacked(affines, axis)
    xs = [_dg._to_device(x, device) for x in Xs]
    if not all(x.stride() == xs[0].stride() for x in xs):
        return None
    if gradient:
        if tuple(out_shapes[0]) != tuple(xs[0].shape):
            raise ValueError("X_shape does not match output shape and cropping.")
        acc = torch.zeros((nb,) + tuple(shp), dtype=xs[0].dtype, device=device)
        ins0, outs0 = acc[0], xs[0]
        in_ptrs = acc.data_ptr() + numpy.arange(nb, dtype=numpy.uint64) * numpy.uint64(acc.stride(0) * acc.element_size())
        out_ptrs = numpy.array([x.data_ptr() for x in xs], dtype=numpy.uint64)
        results = acc
    else:
        res_all = torch.empty((nb,) + tuple(out_shapes[0]), dtype=xs[0].dtype, device=device)
        ins0, outs0 = xs[0], res_all[0]
        in_ptrs = numpy.array([x.data_ptr() for x in xs], dtype=numpy.uint64)
        out_ptrs = res_all.data_ptr() + numpy.arange(nb, dtype=numpy.uint64) * numpy.uint64(res_all.stride(0) * res_all.element_size())
        results = res_all
    disp_ptrs = d_all.data_ptr() + numpy.arange(nb, dtype=numpy.uint64) * numpy.uint64(d_all.stride(0) * d_all.element_size())
    proto, keep = _dg._build_problem([ins0], [outs0], d_all[0], offset, axis, od, md, cv,
                                     None if inv_all is None else inv_all[0], flags)
    u64p = ctypes.POINTER(ctypes.c_uint64)
    in_ptrs = numpy.ascontiguousarray(in_ptrs, dtype=numpy.uint64)
    out_ptrs = numpy.ascontiguousarray(out_ptrs, dtype=numpy.uint64)
    disp_ptrs = numpy.ascontiguousarray(disp_ptrs, dtype=numpy.uint64)
    aff_p = None if inv_all is None else inv_all.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    _lib.check(lib.edf_deform_grid_batch_uniform(ctypes.byref(proto), nb, 1 if gradient else 0,
                                                 in_ptrs.ctypes.data_as(u64p), out_ptrs.ctypes.data_as(u64p),
                                                 disp_ptrs.ctypes.data_as(u64p), aff_p, _dg._stream_ptr(device)))
    return [_dg._from_device(results[b], Xs[b]) for b in range(nb)]


def _prefilter_displacements_stacked(lib, displacements, X0, device):
    """The control grids of the whole batch, prefiltered in naxis launches: grids of one shape and dtype are stacked
    into one (batch, naxis, P...) array and filtered along the grid axes together (the per-line arithmetic is the
    same as for one grid, so the result is bit-identical to ``_prefilter_displacement`` per volume).  Returns a
    tensor indexable by volume, or None when the grids cannot be stacked."""
    torch = _dg.torch
    nb = len(displacements)
    d0 = displacements[0]
    if _dg._is_tensor(d0):
        if not all(_dg._is_tensor(d) and d.shape == d0.shape and d.dtype == d0.dtype and d.device == d0.device
                   for d in displacements):
            return None
        if d0.ndim != X0.ndim + 1 or d0.shape[0] != X0.ndim:
            return None
        stacked = torch.stack([d.detach() for d in displacements]).to(device)
    else:
        arrs = [numpy.asarray(d) for d in displacements]
        if not all(a.shape == arrs[0].shape and a.dtype == arrs[0].dtype for a in arrs):
            return None
        if arrs[0].ndim != X0.ndim + 1 or arrs[0].shape[0] != X0.ndim:
            return None
        stacked = torch.from_numpy(numpy.stack(arrs)).to(device)
    if stacked.dtype not in (torch.float64, torch.float32):
        return None
    out = torch.empty_like(stacked)
    src = stacked
    for ax in range(2, stacked.ndim):
        _dg._spline_filter1d_device(lib, src, out, ax, 3)
        src = out
    return out


def deform_random_grid_batch(Xs, sigma=25, points=3, order=3, mode='constant', cval=0.0, crop=None,
                             prefilter=True, affines=None, generator=None, return_displacements=False):
    """Batched twin of ``deform_random_grid`` (reference deform_grid.py:6-49): every volume of the batch
    gets its own random displacement grid, sampled from N(0, sigma^2) at ``points`` control points per
    axis, and the whole batch goes to the GPU in one ``edf_deform_grid_batch`` call.

    ``generator``: None draws from NumPy's global RNG (``numpy.random.randn``, like the reference, one
    draw per volume in batch order, so ``numpy.random.seed`` reproduces a Python loop over
    ``deform_random_grid``); a ``torch.Generator`` of the CUDA device draws all grids on the device in one
    call (no host round trip; the grids never leave the GPU).  ``return_displacements=True`` also
    returns the grids that were used (a list of arrays / a CUDA tensor of shape (batch, naxis, P...)).
    """
    torch = _dg.torch
    nb = len(Xs)
    if nb == 0:
        return ([], []) if return_displacements else []
    ndim = Xs[0].ndim
    if not isinstance(points, (list, tuple)):
        points = [points] * ndim
    assert len(points) == ndim, 'one number of control points per axis'
    if generator is None:
        Ds = [numpy.random.randn(ndim, *points) * sigma for _ in range(nb)]
    else:
        _dg._require_cuda()
        dev = generator.device
        assert dev.type == 'cuda', 'generator must be a CUDA generator (or None for the NumPy global RNG)'
        Dt = torch.randn((nb, ndim) + tuple(int(q) for q in points), generator=generator, device=dev,
                         dtype=torch.float64) * float(sigma)
        Ds = [Dt[b] for b in range(nb)]
    Ys = deform_grid_batch(Xs, Ds, order=order, mode=mode, cval=cval, crop=crop, prefilter=prefilter,
                           affines=affines)
    if return_displacements:
        return Ys, (Ds if generator is None else Dt)
    return Ys
