#!/usr/bin/env python
"""Where does the host-array (e2e) path spend its time?  Raw PCIe rates of this box next to the
API call in its one-shot and slab-pipelined forms (256^3 float32, order 3, prefilter=False).

    python scripts/e2e_probe.py >> gpurun_out/e2e_probe.jsonl
"""
import sys, os, time, json, importlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import elasticdeform_b200 as edf
dg = importlib.import_module("elasticdeform_b200.deform_grid")

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
N = 256
rng = np.random.default_rng(0)
Xp = torch.empty((N,) * 3, dtype=torch.float32).pin_memory()
Gp = torch.empty((N,) * 3, dtype=torch.float32).pin_memory()
Xp.copy_(torch.from_numpy(rng.random((N,) * 3, dtype=np.float32)))
Gp.copy_(torch.from_numpy(rng.random((N,) * 3, dtype=np.float32)))
Xn, Gn = Xp.numpy(), Gp.numpy()
D = rng.standard_normal((3, 5, 5, 5)) * 8
nbytes = Xp.numel() * 4


def out(**kw):
    print(json.dumps(kw), flush=True)


def timed(fn, reps=8, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3, float(np.min(ts)) * 1e3


# ---- raw PCIe
Xd = torch.empty_like(Xp, device=dev)
Yd = torch.empty_like(Xp, device=dev)
Yp = torch.empty_like(Xp).pin_memory()
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
m, b = timed(lambda: Xd.copy_(Xp, non_blocking=True))
out(what="h2d 67MB pinned", ms=m, best_ms=b, GBps=nbytes / b / 1e6)
m, b = timed(lambda: Yp.copy_(Yd, non_blocking=True))
out(what="d2h 67MB pinned", ms=m, best_ms=b, GBps=nbytes / b / 1e6)


def both():
    with torch.cuda.stream(s1):
        Xd.copy_(Xp, non_blocking=True)
    with torch.cuda.stream(s2):
        Yp.copy_(Yd, non_blocking=True)


m, b = timed(both)
out(what="h2d + d2h concurrently, 67MB each", ms=m, best_ms=b, GBps_each=nbytes / b / 1e6)


def slabbed(k):
    h = N // k
    def f():
        for j in range(k):
            with torch.cuda.stream(s1):
                Xd[j * h:(j + 1) * h].copy_(Xp[j * h:(j + 1) * h], non_blocking=True)
            with torch.cuda.stream(s2):
                Yp[j * h:(j + 1) * h].copy_(Yd[j * h:(j + 1) * h], non_blocking=True)
    return f


for k in (8, 16):
    m, b = timed(slabbed(k))
    out(what="h2d + d2h concurrently in %d slabs" % k, ms=m, best_ms=b)

xpg = torch.from_numpy(np.array(Xn))          # pageable copy
m, b = timed(lambda: Xd.copy_(xpg, non_blocking=True), reps=4, warm=1)
out(what="h2d 67MB pageable", ms=m, best_ms=b, GBps=nbytes / b / 1e6)
t0 = time.perf_counter(); q = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True); t1 = time.perf_counter()
out(what="fresh 67MB pinned allocation", ms=(t1 - t0) * 1e3)
del q

# ---- API calls (results held by the caller across the next call, as a training loop would)
held = {}


def fwd():
    held["y"] = edf.deform_grid(Xn, D, order=3, prefilter=False)


def grad():
    held["dx"] = edf.deform_grid_gradient(Gn, D, order=3, prefilter=False)


def both_calls():
    fwd(); grad()


for slabs in (8, 4, 16):
    dg._PIPELINE_SLABS = slabs
    m, b = timed(fwd); out(what="api forward, pipelined", slabs=slabs, ms=m, best_ms=b)
    m, b = timed(grad); out(what="api gradient, pipelined", slabs=slabs, ms=m, best_ms=b)
dg._PIPELINE_SLABS = 8
m, b = timed(both_calls); out(what="api fwd+grad, pipelined", ms=m, best_ms=b)
keep = dg._PIPELINE_MIN_BYTES
dg._PIPELINE_MIN_BYTES = 1 << 60
m, b = timed(fwd); out(what="api forward, one-shot", ms=m, best_ms=b)
m, b = timed(grad); out(what="api gradient, one-shot", ms=m, best_ms=b)
m, b = timed(both_calls); out(what="api fwd+grad, one-shot", ms=m, best_ms=b)
dg._PIPELINE_MIN_BYTES = keep

# ---- host-side share of one pipelined call: time until the call returns vs GPU-side span
for name, fn in (("forward", fwd), ("gradient", grad)):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    t0 = time.perf_counter()
    fn()
    t1 = time.perf_counter()
    e1.record(); torch.cuda.synchronize()
    out(what="pipelined %s: host wall vs device span" % name, host_ms=(t1 - t0) * 1e3, device_ms=e0.elapsed_time(e1))

