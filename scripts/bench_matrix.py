#!/usr/bin/env python
"""Kernel-time matrix for DESIGN.md: forward / gradient device time of several configurations
(CUDA events, data resident, median of `reps`), with the algorithmic-bytes roofline fraction."""
import ctypes, importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from elasticdeform_b200 import _lib
dg = importlib.import_module("elasticdeform_b200.deform_grid")
lib = _lib.load_library()
dev = torch.device("cuda", 0)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

def run(name, shapes, dtypes, orders, points, sigma, axis=None, mode='constant', reps=20, nsets=3):
    rng = np.random.default_rng(0)
    n = len(shapes)
    naxis = len(points)
    axis_l = [tuple(range(len(s))) if axis is None else axis for s in shapes]
    D = rng.standard_normal((naxis,) + points) * sigma
    d_f = dg._prefilter_displacement(lib, D, dev)
    sets = []
    for k in range(nsets):
        Xs = [torch.from_numpy((rng.random(s) * (5 if 'int' in dt else 1)).astype(dt)).to(dev) for s, dt in zip(shapes, dtypes)]
        Ys = [torch.empty_like(x) for x in Xs]
        dXs = [torch.zeros_like(x) for x in Xs]
        pf, kf = dg._build_problem(Xs, Ys, d_f, None, axis_l, np.array(orders), np.array([dg._MODE_CODES[mode]] * n), np.zeros(n), None)
        pg, kg = dg._build_problem(dXs, Ys, d_f, None, axis_l, np.array(orders), np.array([dg._MODE_CODES[mode]] * n), np.zeros(n), None)
        sets.append((pf, pg, kf, kg, Xs, Ys, dXs))
    st = torch.cuda.current_stream(dev); sp = ctypes.c_void_p(st.cuda_stream)
    out = {}
    for what in ("fwd", "grad"):
        if what == "grad" and any('int' in dt for dt in dtypes) and any(o > 0 for o in orders):
            pass
        ts = []
        for r in range(reps + 3):
            pf, pg, _, _, Xs, Ys, dXs = sets[r % nsets]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            _lib.check((lib.edf_deform_grid if what == "fwd" else lib.edf_deform_grid_grad)(ctypes.byref(pf if what == "fwd" else pg), sp))
            e1.record(st); torch.cuda.synchronize()
            if r >= 3: ts.append(e0.elapsed_time(e1))
        kern = _lib.last_kernel()
        t = float(np.median(ts))
        nbytes = sum(2 * int(np.prod(s)) * np.dtype(dt).itemsize for s, dt in zip(shapes, dtypes))
        nvox = int(np.prod([shapes[0][a] for a in axis_l[0]]))
        out[what] = dict(ms=round(t, 4), kernel=kern, GBps=round(nbytes / t / 1e6, 1), frac=round(nbytes / t / 1e6 / PEAK, 4),
                         Mvox_s=round(nvox / t / 1e3, 1))
    print(json.dumps({"config": name, **out}), flush=True)

def cfg4_batch(nb=64, reps=5):
    """BASELINE config 4 on one GPU: nb volumes 128^3 -> crop 64^3, per-volume displacement and 3-D affine
    (rotation 15 deg about axis 0, zoom 1.2 about the crop centre), order 3, one edf_deform_grid_batch call."""
    from elasticdeform_b200 import batch
    rng = np.random.default_rng(4)
    th = np.radians(15.0)
    R = np.array([[1, 0, 0], [0, np.cos(th), -np.sin(th)], [0, np.sin(th), np.cos(th)]]) * 1.2
    c = np.array([31.5, 31.5, 31.5])
    A = np.concatenate([R, (c - R @ c)[:, None]], axis=1)
    Xs = [torch.from_numpy(rng.random((128,) * 3, dtype=np.float32)).to(dev) for _ in range(nb)]
    Ds = [rng.standard_normal((3, 5, 5, 5)) * 8.0 for _ in range(nb)]
    crop = (slice(32, 96),) * 3
    st = torch.cuda.current_stream(dev)
    out = {}
    for what in ("fwd", "grad"):
        ts = []
        src = Xs
        if what == "grad":
            src = [torch.from_numpy(rng.random((64,) * 3, dtype=np.float32)).to(dev) for _ in range(nb)]
        for r in range(reps + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(st)
            batch.deform_grid_batch(src, Ds, order=3, crop=crop, prefilter=False, affines=[A] * nb,
                                    gradient=(what == "grad"), X_shape=(128,) * 3)
            e1.record(st); torch.cuda.synchronize()
            if r >= 2: ts.append(e0.elapsed_time(e1))
        t = float(np.median(ts))
        out[what] = dict(ms=round(t, 3), Mvox_out_s=round(nb * 64 ** 3 / t / 1e3, 1))
    print(json.dumps({"config": "cfg4 batch %dx(128^3 -> crop 64^3, affine), order 3, incl. host prep" % nb, **out}), flush=True)


def prefilter_times(shape=(256, 256, 256), order=3, reps=5):
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.random(shape, dtype=np.float32)).to(dev)
    out = torch.empty_like(x)
    st = torch.cuda.current_stream(dev)
    res = {}
    for adj in (False, True):
        for ax in range(len(shape)):
            ts = []
            for r in range(reps + 2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                dg._spline_filter1d_device(lib, x, out, ax, order, adjoint=adj)
                e1.record(st); torch.cuda.synchronize()
                if r >= 2: ts.append(e0.elapsed_time(e1))
            res["%s_axis%d_ms" % ("adjoint" if adj else "prefilter", ax)] = round(float(np.median(ts)), 4)
    print(json.dumps({"config": "prefilter %s order %d" % (str(shape), order), **res}), flush=True)


if __name__ == "__main__":
    S = (256, 256, 256)
    for o in (0, 1, 2, 3):
        run("256^3 f32 order %d" % o, [S], ['float32'], [o], (5, 5, 5), 8.0)
    run("256^3 f64 order 3 (fast_f64)", [S], ['float64'], [3], (5, 5, 5), 8.0, reps=10)
    run("cfg3 256^3 f32 o3 + int32 o0", [S, S], ['float32', 'int32'], [3, 0], (5, 5, 5), 8.0)
    run("cfg2 128^3 f32 o3", [(128,) * 3], ['float32'], [3], (5, 5, 5), 8.0)
    run("cfg5 32x128^3 f32 o1 axis=(1,2,3)", [(32, 128, 128, 128)], ['float32'], [1], (5, 5, 5), 8.0, axis=(1, 2, 3), reps=10)
    run("cfg1 200x300 f32 o3 reflect", [(200, 300)], ['float32'], [3], (3, 3), 25.0, mode='reflect')
    run("256^3 f32 order 3 mirror", [S], ['float32'], [3], (5, 5, 5), 8.0, mode='mirror', reps=10)
    run("256^3 f32 order 1 nearest", [S], ['float32'], [1], (5, 5, 5), 8.0, mode='nearest', reps=10)
    prefilter_times()
    cfg4_batch()
