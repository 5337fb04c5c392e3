"""ctypes binding of the C-ABI library (include/edf_b200.h).

This is the stub a maintainer of the reference would add in place of
``from . import _deform_grid`` (reference deform_grid.py:4): three calls,
``deform_grid`` / ``deform_grid_grad`` / ``spline_filter1d_grad`` (reference
_deform_grid.c:306-311), plus the forward prefilter that the reference borrows
from SciPy.  There is no CPU fallback: if the library is missing or no B200 is
visible every call raises.
"""
import ctypes
import os
import threading

import numpy

EDF_MAX_DIMS = 8
EDF_MAX_AXIS = 4
EDF_MAX_INPUTS = 8

EDF_FLAG_FORCE_GENERIC = 1
EDF_FLAG_NO_WINDOW = 2
EDF_FLAG_STAGED_FWD = 4
EDF_FLAG_FIXED_WINDOW = 8
EDF_FLAG_STAGED_ALL = 16
EDF_FLAG_STEEP = 32

# numpy dtype -> edf_dtype (the 11 distinct element types of deform.c:863-888)
_DTYPE_CODES = {
    numpy.dtype('bool'): 0, numpy.dtype('uint8'): 1, numpy.dtype('uint16'): 2,
    numpy.dtype('uint32'): 3, numpy.dtype('uint64'): 4, numpy.dtype('int8'): 5,
    numpy.dtype('int16'): 6, numpy.dtype('int32'): 7, numpy.dtype('int64'): 8,
    numpy.dtype('float32'): 9, numpy.dtype('float64'): 10,
}


def dtype_code(np_dtype):
    try:
        return _DTYPE_CODES[numpy.dtype(np_dtype)]
    except (KeyError, TypeError):
        # same exception type and text as the reference (deform.c:889-893)
        raise RuntimeError('data type not supported')


class EdfArray(ctypes.Structure):
    _fields_ = [
        ("data", ctypes.c_void_p),
        ("dtype", ctypes.c_int32),
        ("ndim", ctypes.c_int32),
        ("shape", ctypes.c_int64 * EDF_MAX_DIMS),
        ("strides", ctypes.c_int64 * EDF_MAX_DIMS),
    ]


class EdfProblem(ctypes.Structure):
    _fields_ = [
        ("ninputs", ctypes.c_int32),
        ("naxis", ctypes.c_int32),
        ("inputs", ctypes.POINTER(EdfArray)),
        ("outputs", ctypes.POINTER(EdfArray)),
        ("displacement", EdfArray),
        ("output_offset", ctypes.POINTER(ctypes.c_int64)),
        ("axis", ctypes.POINTER(ctypes.c_int32)),
        ("orders", ctypes.POINTER(ctypes.c_int32)),
        ("modes", ctypes.POINTER(ctypes.c_int32)),
        ("cvals", ctypes.POINTER(ctypes.c_double)),
        ("affine", ctypes.POINTER(ctypes.c_double)),
        ("flags", ctypes.c_uint32),
    ]


# every symbol include/edf_b200.h declares (checked by tests/test_host_api.py)
EXPORTED_SYMBOLS = (
    "edf_deform_grid", "edf_deform_grid_grad", "edf_deform_grid_batch", "edf_deform_grid_batch_uniform",
    "edf_spline_filter1d", "edf_spline_filter1d_grad", "edf_last_error",
    "edf_version", "edf_device_ok", "edf_launch_count", "edf_last_kernel", "edf_debug_tile_profile",
)

_lib = None
_lib_lock = threading.Lock()


def library_path():
    # EDF_B200_LIB: another build of the same library (kernel A/B experiments, scripts/ab_variants.sh)
    return os.environ.get("EDF_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libedf_b200.so")


def load_library():
    """Load libedf_b200.so (once). Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                "elasticdeform_b200: CUDA library %s is missing. Build it with "
                "`python -m elasticdeform_b200.build` (needs nvcc). There is no CPU fallback." % path)
        lib = ctypes.CDLL(path)
        vp, i32 = ctypes.c_void_p, ctypes.c_int32
        lib.edf_deform_grid.argtypes = [ctypes.POINTER(EdfProblem), vp]
        lib.edf_deform_grid.restype = ctypes.c_int
        lib.edf_deform_grid_grad.argtypes = [ctypes.POINTER(EdfProblem), vp]
        lib.edf_deform_grid_grad.restype = ctypes.c_int
        lib.edf_deform_grid_batch.argtypes = [ctypes.POINTER(EdfProblem), i32, i32, vp]
        lib.edf_deform_grid_batch.restype = ctypes.c_int
        u64p = ctypes.POINTER(ctypes.c_uint64)
        lib.edf_deform_grid_batch_uniform.argtypes = [ctypes.POINTER(EdfProblem), i32, i32, u64p, u64p, u64p,
                                                      ctypes.POINTER(ctypes.c_double), vp]
        lib.edf_deform_grid_batch_uniform.restype = ctypes.c_int
        lib.edf_spline_filter1d.argtypes = [ctypes.POINTER(EdfArray), ctypes.POINTER(EdfArray), i32, i32, vp]
        lib.edf_spline_filter1d.restype = ctypes.c_int
        lib.edf_spline_filter1d_grad.argtypes = [ctypes.POINTER(EdfArray), ctypes.POINTER(EdfArray), i32, i32, vp]
        lib.edf_spline_filter1d_grad.restype = ctypes.c_int
        lib.edf_last_error.restype = ctypes.c_char_p
        lib.edf_last_kernel.restype = ctypes.c_char_p
        lib.edf_version.restype = ctypes.c_int
        lib.edf_device_ok.restype = ctypes.c_int
        lib.edf_launch_count.restype = ctypes.c_uint64
        _lib = lib
    return _lib


def check(rc):
    """Map edf_status codes to the exception types the reference raises."""
    if rc == 0:
        return
    msg = load_library().edf_last_error().decode("utf-8", "replace")
    if rc == -2:
        raise ValueError(msg)
    if rc == -3:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def make_array(data_ptr, np_dtype, shape, strides_bytes):
    a = EdfArray()
    nd = len(shape)
    if nd > EDF_MAX_DIMS:
        raise RuntimeError('too many dimensions (max %d)' % EDF_MAX_DIMS)
    a.data = data_ptr
    a.dtype = dtype_code(np_dtype)
    a.ndim = nd
    for i in range(nd):
        a.shape[i] = int(shape[i])
        a.strides[i] = int(strides_bytes[i])
    return a


def launch_count():
    return int(load_library().edf_launch_count())


def last_kernel():
    return load_library().edf_last_kernel().decode()
