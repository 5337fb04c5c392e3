"""CPU oracle for the elastic-deformation path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Nothing under elasticdeform_b200/ does.

Two interchangeable back ends behind the same NumPy-level functions:

* ``impl='ref'``  -- the UNMODIFIED reference C extension compiled from
  /root/reference by oracle/build_ref.py into oracle/_ref/ (kind "reference"),
  with SciPy doing the prefilter exactly as reference deform_grid.py:155-169;
* ``impl='port'`` -- oracle/deform_oracle.c, a plain-C restatement (kind "port"),
  including its own restatement of SciPy's prefilter.

The NumPy-level glue below restates the host preparation of the reference
(deform_grid.py:52-291; cited per function) in the oracle's own words so that the
product's host code (elasticdeform_b200/deform_grid.py) is checked against an
independent implementation.  Parity status: pinned (tests/test_oracle.py checks
port == ref bit for bit, and both against tests/golden/*.npz generated from the
reference package itself).
"""
import ctypes
import importlib.util
import os
import subprocess
import sysconfig

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
PORT_SO = os.path.join(BUILD_DIR, "liboracle.so")
PORT_SRC = os.path.join(HERE, "deform_oracle.c")

MODES = {'nearest': 0, 'wrap': 1, 'reflect': 2, 'mirror': 3, 'constant': 4}
_DT = {np.dtype(k): v for k, v in [
    ('bool', 0), ('uint8', 1), ('uint16', 2), ('uint32', 3), ('uint64', 4), ('int8', 5),
    ('int16', 6), ('int32', 7), ('int64', 8), ('float32', 9), ('float64', 10)]}
ORC_MAXDIMS = 16


# ----------------------------------------------------------------------------------------------
# back ends
# ----------------------------------------------------------------------------------------------
def build_port(force=False):
    """gcc -O2 -ffp-contract=off oracle/deform_oracle.c -> oracle/_build/liboracle.so"""
    if (not force and os.path.exists(PORT_SO)
            and os.path.getmtime(PORT_SO) >= os.path.getmtime(PORT_SRC)):
        return PORT_SO
    os.makedirs(BUILD_DIR, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", PORT_SO,
                           PORT_SRC, "-lm"])
    return PORT_SO


class _OrcArray(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("dtype", ctypes.c_int32), ("ndim", ctypes.c_int32),
                ("shape", ctypes.c_int64 * ORC_MAXDIMS), ("strides", ctypes.c_int64 * ORC_MAXDIMS)]


def _orc(a):
    s = _OrcArray()
    s.data = a.ctypes.data
    s.dtype = _DT[a.dtype]
    s.ndim = a.ndim
    for i in range(a.ndim):
        s.shape[i] = a.shape[i]
        s.strides[i] = a.strides[i]
    return s


_port = None


def port_lib():
    global _port
    if _port is None:
        lib = ctypes.CDLL(build_port())
        lib.orc_deform_grid.restype = ctypes.c_int
        lib.orc_spline_filter1d.restype = ctypes.c_int
        lib.orc_spline_filter1d_grad.restype = ctypes.c_int
        _port = lib
    return _port


class _PortModule(object):
    """Same three entry points and argument conventions as the reference extension module
    (_deform_grid.c:306-311)."""

    @staticmethod
    def _call(gradient, inputs, displacement, output_offset, outputs, axis, orders, modes, cvals, affine):
        lib = port_lib()
        n = len(inputs)
        naxis = len(axis[0])
        ins = (_OrcArray * n)(*[_orc(a) for a in inputs])
        outs = (_OrcArray * n)(*[_orc(a) for a in outputs])
        d = _orc(displacement)
        off = None
        if output_offset is not None:
            off = (ctypes.c_int64 * naxis)(*[int(v) for v in output_offset])
        ax = (ctypes.c_int * (n * naxis))(*[int(a) for t in axis for a in t])
        od = (ctypes.c_int * n)(*[int(v) for v in orders])
        md = (ctypes.c_int * n)(*[int(v) for v in modes])
        cv = (ctypes.c_double * n)(*[float(v) for v in cvals])
        af = None
        if affine is not None:
            flat = np.ascontiguousarray(affine, dtype='float64').ravel()
            af = (ctypes.c_double * flat.size)(*flat.tolist())
        rc = lib.orc_deform_grid(int(gradient), n, ins, ctypes.byref(d), off, outs, naxis, ax, od, md, cv, af)
        if rc:
            raise RuntimeError("oracle port: unsupported arguments")

    def deform_grid(self, *a):
        self._call(0, *a)

    def deform_grid_grad(self, *a):
        self._call(1, *a)

    @staticmethod
    def spline_filter1d_grad(inp, out, axis, order):
        a, b = _orc(inp), _orc(out)
        if port_lib().orc_spline_filter1d_grad(ctypes.byref(a), ctypes.byref(b), int(axis), int(order)):
            raise RuntimeError("oracle port: unsupported arguments")

    @staticmethod
    def spline_filter1d(inp, axis, order, output):
        a, b = _orc(inp), _orc(output)
        if port_lib().orc_spline_filter1d(ctypes.byref(a), ctypes.byref(b), int(axis), int(order)):
            raise RuntimeError("oracle port: unsupported arguments")


_ref = None


def ref_so_path():
    return os.path.join(HERE, "_ref", "_deform_grid" + sysconfig.get_config_var("EXT_SUFFIX"))


def ref_available():
    return os.path.exists(ref_so_path())


def ref_module():
    """The compiled, unmodified reference extension (oracle/_ref)."""
    global _ref
    if _ref is None:
        path = ref_so_path()
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref is not built: run `python oracle/build_ref.py` where "
                               "/root/reference exists")
        spec = importlib.util.spec_from_file_location("_deform_grid", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref = mod
    return _ref


_extra_backends = {}


def register_backend(name, module, prefilter):
    """Let a test plug another implementation of the three C entry points (e.g. the host
    build of the device code, tests/hostsim.py) under the same NumPy-level glue."""
    _extra_backends[name] = (module, prefilter)


def _backend(impl):
    if impl in _extra_backends:
        return _extra_backends[impl]
    if impl == 'ref':
        import scipy.ndimage
        mod = ref_module()

        def prefilter(x, axis, order, output):
            scipy.ndimage.spline_filter1d(x, axis=axis, order=order, output=output)
        return mod, prefilter
    if impl == 'port':
        mod = _PortModule()
        return mod, mod.spline_filter1d
    raise ValueError(impl)


# ----------------------------------------------------------------------------------------------
# NumPy-level restatement of the reference's host preparation
# ----------------------------------------------------------------------------------------------
def _as_list(v, n):
    return list(v) if isinstance(v, (list, tuple)) else [v] * n


def _axes(axis, Xs):                      # deform_grid.py:308-326
    if axis is None:
        axis = [tuple(range(x.ndim)) for x in Xs]
    elif isinstance(axis, int):
        axis = (axis,)
    if isinstance(axis, tuple):
        axis = [axis] * len(Xs)
    shapes = {tuple(x.shape[d] for d in ax) for x, ax in zip(Xs, axis)}
    assert len(shapes) == 1
    return list(axis), shapes.pop()


def _crop(shapes_in, axis, deform_shape, crop):   # deform_grid.py:328-354
    if crop is None:
        return [tuple(s) for s in shapes_in], None
    shapes = [list(s) for s in shapes_in]
    offset = [0] * len(deform_shape)
    for d, sl in enumerate(crop):
        start = sl.start or 0
        stop = sl.stop or deform_shape[d]
        for i in range(len(shapes)):
            shapes[i][axis[i][d]] = stop - start
        offset[d] = start
    off = np.array(offset).astype('int64') if any(o > 0 for o in offset) else None
    return [tuple(s) for s in shapes], off


def _inverse_affine(affine, naxis):       # deform_grid.py:381-399
    if affine is None:
        return None
    affine = np.asarray(affine)
    if affine.shape == (naxis + 1, naxis + 1):
        affine = affine[:naxis, :]
    affine = np.array(affine).astype('float64')
    inv = np.zeros(affine.shape, dtype='float64')
    inv[:, :-1] = np.linalg.inv(affine[:, :-1])
    inv[:, -1] = -np.dot(inv[:, :-1], affine[:, -1])
    return inv


def _rot_zoom(rotate, zoom, inv, out_shape):      # deform_grid.py:401-438
    if rotate is None and zoom is None:
        return inv
    assert len(out_shape) == 2
    angle = -float(rotate or 0)
    z = 1 / float(zoom or 1)
    c = np.array(out_shape) / 2 - 0.5
    m = np.array([[1, 0, -c[0]], [0, 1, -c[1]], [0, 0, 1]])
    if angle:
        th = np.radians(angle)
        m = np.dot(np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]]), m)
    if z:
        m = np.dot(np.array([[z, 0, 0], [0, z, 0], [0, 0, 1]]), m)
    m = np.dot(np.array([[1, 0, c[0]], [0, 1, c[1]], [0, 0, 1]]), m)
    if inv is not None:
        base = np.eye(3, dtype='float64')
        base[:-1, :] = inv
        return np.dot(m, base)[:2, :]
    return m[:2, :]


def _prefilter_displacement(displacement, prefilter_fn):    # deform_grid.py:166-169
    d_f = np.zeros_like(displacement)
    for d in range(1, displacement.ndim):
        prefilter_fn(displacement, d, 3, d_f)
        displacement = d_f
    return d_f


def deform_grid(X, displacement, order=3, mode='constant', cval=0.0, crop=None, prefilter=True,
                axis=None, affine=None, rotate=None, zoom=None, impl='port'):
    """Oracle forward pass, reference deform_grid.py:52-179."""
    mod, pf = _backend(impl)
    Xs = X if isinstance(X, list) else [X]
    n = len(Xs)
    axis, deform_shape = _axes(axis, Xs)
    out_shapes, offset = _crop([x.shape for x in Xs], axis, deform_shape, crop)
    order = np.array(_as_list(order, n)).astype('int64')
    mode = np.array([MODES[m] for m in _as_list(mode, n)]).astype('int64')
    cval = np.array(_as_list(cval, n)).astype('float64')
    inv = _inverse_affine(affine, len(axis[0]))
    inv = _rot_zoom(rotate, zoom, inv, [out_shapes[0][d] for d in axis[0]])
    Xs_f = []
    for i, x in enumerate(Xs):
        if prefilter and order[i] > 1:
            x_f = np.zeros_like(x)
            for d in axis[i]:
                pf(x, d, int(order[i]), x_f)
                x = x_f
            Xs_f.append(x_f)
        else:
            Xs_f.append(x)
    d_f = _prefilter_displacement(displacement, pf)
    outputs = [np.zeros(s, dtype=x.dtype) for s, x in zip(out_shapes, Xs)]
    mod.deform_grid(Xs_f, d_f, offset, outputs, axis, order, mode, cval, inv)
    return outputs if isinstance(X, list) else outputs[0]


def deform_grid_gradient(dY, displacement, order=3, mode='constant', cval=0.0, crop=None,
                         prefilter=True, axis=None, X_shape=None, affine=None, rotate=None,
                         zoom=None, impl='port'):
    """Oracle backward pass, reference deform_grid.py:182-291."""
    mod, pf = _backend(impl)
    dYs = dY if isinstance(dY, list) else [dY]
    n = len(dYs)
    if isinstance(X_shape, tuple):
        X_shape = [X_shape]
    elif X_shape is None:
        X_shape = [dy.shape for dy in dYs]
    dXs = [np.zeros(s, dy.dtype) for s, dy in zip(X_shape, dYs)]
    axis, deform_shape = _axes(axis, dXs)
    out_shapes, offset = _crop([x.shape for x in dXs], axis, deform_shape, crop)
    assert [tuple(s) for s in out_shapes] == [dy.shape for dy in dYs]
    order = np.array(_as_list(order, n)).astype('int64')
    mode = np.array([MODES[m] for m in _as_list(mode, n)]).astype('int64')
    cval = np.array(_as_list(cval, n)).astype('float64')
    inv = _inverse_affine(affine, len(axis[0]))
    inv = _rot_zoom(rotate, zoom, inv, [out_shapes[0][d] for d in axis[0]])
    d_f = _prefilter_displacement(displacement, pf)
    mod.deform_grid_grad(dXs, d_f, offset, dYs, axis, order, mode, cval, inv)
    res = []
    for i, x in enumerate(dXs):
        if prefilter and order[i] > 1:
            x_f = np.zeros_like(x)
            for d in axis[i]:
                mod.spline_filter1d_grad(x, x_f, d, int(order[i]))
                x = x_f
            res.append(x_f)
        else:
            res.append(x)
    return res if isinstance(dY, list) else res[0]


def spline_filter1d(x, axis, order, impl='port'):
    """Forward prefilter along one axis into a new array of x's dtype."""
    _, pf = _backend(impl)
    out = np.zeros_like(x)
    pf(x, axis, order, out)
    return out


def spline_filter1d_grad(x, axis, order, impl='port'):
    mod, _ = _backend(impl)
    out = np.zeros_like(x)
    mod.spline_filter1d_grad(x, out, axis, order)
    return out
