"""TEST-ONLY host build of the device code (tests/_hostsim/hostsim.cpp).

Compiles the __host__ __device__ per-voxel routines the sm_100a kernels execute
(elasticdeform_b200/csrc/edf_core.h, edf_spline_lines.h, edf_fast_core.h) with g++
so that `pytest -m "not gpu"` can check them against the oracle on a CPU-only box.
Never imported by the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

from elasticdeform_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "_hostsim", "hostsim.cpp")
OUT_DIR = os.path.join(HERE, "_hostsim", "_build")
SO = os.path.join(OUT_DIR, "libedf_hostsim.so")
CSRC = os.path.join(os.path.dirname(HERE), "elasticdeform_b200", "csrc")


def build(force=False):
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                           "-o", SO, SRC, "-lm"])
    return SO


_so = None


def lib():
    global _so
    if _so is None:
        so = ctypes.CDLL(build())
        so.hostsim_last_error.restype = ctypes.c_char_p
        _so = so
    return _so


def _arr(a):
    return _lib.make_array(a.ctypes.data if a.size else 0, a.dtype, a.shape, a.strides)


def _problem(inputs, displacement, output_offset, outputs, axis, orders, modes, cvals, affine):
    n = len(inputs)
    naxis = len(axis[0])
    keep = []
    ins = (_lib.EdfArray * n)(*[_arr(a) for a in inputs])
    outs = (_lib.EdfArray * n)(*[_arr(a) for a in outputs])
    pr = _lib.EdfProblem()
    pr.ninputs, pr.naxis = n, naxis
    pr.inputs, pr.outputs = ins, outs
    pr.displacement = _arr(displacement)
    if output_offset is not None:
        off = (ctypes.c_int64 * naxis)(*[int(v) for v in output_offset])
        pr.output_offset = off
        keep.append(off)
    ax = (ctypes.c_int32 * (n * naxis))(*[int(a) for t in axis for a in t])
    od = (ctypes.c_int32 * n)(*[int(v) for v in orders])
    md = (ctypes.c_int32 * n)(*[int(v) for v in modes])
    cv = (ctypes.c_double * n)(*[float(v) for v in cvals])
    pr.axis, pr.orders, pr.modes, pr.cvals = ax, od, md, cv
    keep += [ins, outs, ax, od, md, cv]
    if affine is not None:
        flat = np.ascontiguousarray(affine, dtype='float64').ravel()
        af = (ctypes.c_double * flat.size)(*flat.tolist())
        pr.affine = af
        keep.append(af)
    return pr, keep


def _check(rc):
    if rc:
        raise RuntimeError(lib().hostsim_last_error().decode())


class HostSimModule(object):
    """Entry points with the reference extension's argument conventions
    (_deform_grid.c:306-311), backed by the host-compiled device code."""

    def deform_grid(self, inputs, displacement, output_offset, outputs, axis, orders, modes, cvals, affine):
        d = np.ascontiguousarray(displacement, dtype='float64') if displacement.dtype not in (np.float32, np.float64) else displacement
        pr, keep = _problem(inputs, d, output_offset, outputs, axis, orders, modes, cvals, affine)
        _check(lib().hostsim_deform(ctypes.byref(pr), 0))

    def deform_grid_grad(self, inputs, displacement, output_offset, outputs, axis, orders, modes, cvals, affine):
        d = np.ascontiguousarray(displacement, dtype='float64') if displacement.dtype not in (np.float32, np.float64) else displacement
        pr, keep = _problem(inputs, d, output_offset, outputs, axis, orders, modes, cvals, affine)
        _check(lib().hostsim_deform(ctypes.byref(pr), 1))

    def spline_filter1d_grad(self, inp, out, axis, order):
        a, b = _arr(inp), _arr(out)
        _check(lib().hostsim_filter(ctypes.byref(a), ctypes.byref(b), int(axis), int(order), 1))

    def spline_filter1d(self, inp, axis, order, output):
        a, b = _arr(inp), _arr(output)
        _check(lib().hostsim_filter(ctypes.byref(a), ctypes.byref(b), int(axis), int(order), 0))


def fast_coords(inputs, displacement, output_offset, outputs, axis, orders, modes, cvals, affine, ii=0,
                exact=False):
    """Window starts / fractional offsets / constant flags the fast kernels derive
    (exact=True: the same quantities from the reference-order evaluation of every voxel)."""
    pr, keep = _problem(inputs, displacement, output_offset, outputs, axis, orders, modes, cvals, affine)
    naxis = len(axis[0])
    nvox = int(np.prod([outputs[0].shape[a] for a in axis[0]]))
    starts = np.zeros((nvox, naxis), dtype=np.int64)
    fracs = np.zeros((nvox, naxis), dtype=np.float32)
    const = np.zeros(nvox, dtype=np.uint8)
    nex = ctypes.c_int64(0)
    if exact:
        _check(lib().hostsim_exact_coords(ctypes.byref(pr), int(ii), starts.ctypes.data_as(ctypes.c_void_p),
                                          fracs.ctypes.data_as(ctypes.c_void_p),
                                          const.ctypes.data_as(ctypes.c_void_p)))
        return starts, fracs, const, 0
    _check(lib().hostsim_fast_coords(ctypes.byref(pr), int(ii), starts.ctypes.data_as(ctypes.c_void_p),
                                     fracs.ctypes.data_as(ctypes.c_void_p),
                                     const.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nex)))
    return starts, fracs, const, nex.value
