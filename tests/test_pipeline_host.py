"""Host-side choreography of the slab pipeline (elasticdeform_b200/deform_grid.py) without a GPU: the CUDA
streams / events are replaced by inert stand-ins, device memory by host memory, and the C-ABI library by a
recorder that checks every launch it is handed.  What is verified here is the Python logic only -- which
slabs are launched with which crop offset / output pointer / extent, which prefilter passes run on which
slab views and when, what is copied back -- not a single voxel value (tests/test_parity_gpu.py does that on
the B200)."""
import contextlib
import ctypes
import importlib

import numpy as np
import pytest
import scipy.ndimage as ndi
import torch

dg = importlib.import_module("elasticdeform_b200.deform_grid")
from elasticdeform_b200 import _lib


class _Stream(object):
    cuda_stream = 0

    def wait_stream(self, other): pass
    def wait_event(self, ev): pass
    def synchronize(self): pass


class _Event(object):
    def __init__(self, enable_timing=False): pass
    def record(self, stream=None): pass


class _Recorder(object):
    """Stands in for libedf_b200.so: records launches and filter calls."""

    def __init__(self):
        self.launches, self.filters = [], []

    def _launch(self, gradient, ref, stream):
        pr = ctypes.cast(ref, ctypes.POINTER(_lib.EdfProblem)).contents
        outs = [(pr.outputs[i].data, pr.outputs[i].shape[0], pr.outputs[i].strides[0]) for i in range(pr.ninputs)]
        ins = [(pr.inputs[i].data, pr.inputs[i].shape[0]) for i in range(pr.ninputs)]
        self.launches.append(dict(gradient=gradient, off0=pr.output_offset[0], outs=outs, ins=ins,
                                  filters_before=len(self.filters)))
        return 0

    def edf_deform_grid(self, ref, stream): return self._launch(0, ref, stream)
    def edf_deform_grid_grad(self, ref, stream): return self._launch(1, ref, stream)

    def _filter(self, adjoint, a_in, a_out, axis, order, stream):
        ai = ctypes.cast(a_in, ctypes.POINTER(_lib.EdfArray)).contents
        ao = ctypes.cast(a_out, ctypes.POINTER(_lib.EdfArray)).contents
        self.filters.append(dict(adjoint=adjoint, axis=axis, order=order, src=ai.data, dst=ao.data, planes=ai.shape[0],
                                 launches_before=len(self.launches)))
        if ai.data != ao.data:                       # stand-in for the filter: copy (the arrays here are contiguous)
            ctypes.memmove(ao.data, ai.data, ai.shape[0] * ai.strides[0])
        return 0

    def edf_spline_filter1d(self, *a): return self._filter(False, *a)
    def edf_spline_filter1d_grad(self, *a): return self._filter(True, *a)


@pytest.fixture
def host_pipeline(monkeypatch):
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: _Stream())
    monkeypatch.setattr(torch.cuda, "Stream", lambda device=None: _Stream())
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(dg, "_SIDE_STREAMS", {})

    def prefiltered(lib, displacement, device):
        c = np.array(displacement, dtype=np.float64)
        for a in range(1, c.ndim):
            if c.shape[a] > 1:
                c = ndi.spline_filter1d(c, 3, axis=a, mode='mirror')
        return torch.from_numpy(c)

    monkeypatch.setattr(dg, "_prefilter_displacement", prefiltered)
    return _Recorder()


def _slabs(n, h):
    return [(a, min(n, a + h)) for a in range(0, n, h)]


@pytest.mark.parametrize("prefilter,order,crop0", [(False, 3, None), (True, 3, None), (True, 1, (6, 58)), (True, 2, (6, 58))])
def test_forward_choreography(host_pipeline, prefilter, order, crop0):
    rec, dev = host_pipeline, torch.device("cpu")
    X = np.arange(64 * 12 * 10, dtype=np.float32).reshape(64, 12, 10)
    D = np.random.default_rng(0).standard_normal((3, 4, 3, 3)) * 4
    off0, out0 = (0, 64) if crop0 is None else (crop0[0], crop0[1] - crop0[0])
    h = 8
    offset = None if crop0 is None else np.array([off0, 0, 0])
    res = dg._pipelined_forward(rec, dev, [X], D, [(out0, 12, 10)], offset, [(0, 1, 2)], np.array([order]),
                                np.array([4]), np.array([0.0]), h, 0, prefilter)
    assert res is not None and res[0].shape == (out0, 12, 10) and res[0].dtype == np.float32
    slabs = _slabs(out0, h)
    assert len(rec.launches) == len(slabs)
    base, step = rec.launches[0]["outs"][0][0], rec.launches[0]["outs"][0][2]
    for (a, b), L in zip(slabs, rec.launches):
        assert L["gradient"] == 0 and L["off0"] == off0 + a
        assert L["outs"][0][1] == b - a and L["outs"][0][0] == base + a * step       # output slab [a, b)
        assert L["ins"][0][1] == 64                                                   # the whole input volume
    pf = prefilter and order > 1
    if not pf:
        assert rec.filters == []
        return
    # axis 0 once over the whole volume before any launch, then axes 1, 2 per input slab, in place, each
    # slab complete before the first launch that may read it
    f0 = rec.filters[0]
    assert f0["axis"] == 0 and f0["planes"] == 64 and f0["launches_before"] == 0 and not f0["adjoint"]
    rest = rec.filters[1:]
    assert [f["axis"] for f in rest] == [1, 2] * 8 and all(f["src"] == f["dst"] and f["planes"] == h for f in rest)
    fbase = f0["dst"]
    assert [f["src"] for f in rest[0::2]] == [fbase + j * h * 12 * 10 * 4 for j in range(8)]
    reach = dg._slab_reach(dg._prefilter_displacement(None, D, dev), np.array([order]), 64, off0, slabs)
    for (a, b), L, (lo, hi) in zip(slabs, rec.launches, reach):
        need = min(63, max(0, b - 1 + off0 + hi)) // h
        assert (L["filters_before"] - 1) // 2 >= need + 1, "launch before its input slabs were prefiltered"
        assert L["ins"][0][0] == fbase                                                # gathers from the filtered copy


@pytest.mark.parametrize("prefilter,order", [(False, 3), (True, 3), (True, 1)])
def test_gradient_choreography(host_pipeline, prefilter, order):
    rec, dev = host_pipeline, torch.device("cpu")
    G = np.ones((64, 12, 10), dtype=np.float32)
    D = np.random.default_rng(1).standard_normal((3, 4, 3, 3)) * 4
    h = 8
    res = dg._pipelined_gradient(rec, dev, [G], [(64, 12, 10)], D, None, [(0, 1, 2)], np.array([order]),
                                 np.array([4]), np.array([0.0]), h, 0, prefilter)
    assert res is not None and res[0].shape == (64, 12, 10)
    slabs = _slabs(64, h)
    assert len(rec.launches) == len(slabs)
    base, step = rec.launches[0]["outs"][0][0], rec.launches[0]["outs"][0][2]
    for (a, b), L in zip(slabs, rec.launches):
        assert L["gradient"] == 1 and L["off0"] == a and L["outs"][0][1] == b - a and L["outs"][0][0] == base + a * step
        assert L["ins"][0][1] == 64                                                   # dX: the whole volume
    if prefilter and order > 1:
        # adjoint passes only after the last scatter: axis 0 whole, then axes 1, 2 per slab
        assert all(f["adjoint"] and f["launches_before"] == len(slabs) for f in rec.filters)
        assert rec.filters[0]["axis"] == 0 and rec.filters[0]["planes"] == 64
        assert [f["axis"] for f in rec.filters[1:]] == [1, 2] * 8
    else:
        assert rec.filters == []
    # the result buffer received every plane: the stand-in device memory is host memory, dX starts as zeros and
    # nothing scattered into it, so the copy-back must have overwritten the (uninitialised) result with zeros
    assert float(np.abs(res[0]).max()) == 0.0


def test_non_finite_displacement_falls_back(host_pipeline):
    rec, dev = host_pipeline, torch.device("cpu")
    X = np.zeros((64, 12, 10), dtype=np.float32)
    D = np.zeros((3, 3, 3, 3))
    D[0, 1, 1, 1] = np.inf
    res = dg._pipelined_forward(rec, dev, [X], D, [(64, 12, 10)], None, [(0, 1, 2)], np.array([1]),
                                np.array([4]), np.array([0.0]), 8, 0, False)
    assert res is None and rec.launches == []
