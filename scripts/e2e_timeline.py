#!/usr/bin/env python
"""Device-side timeline of one slab-pipelined API call (256^3 float32, order 3, prefilter=False): when
each upload, kernel and download finishes relative to the start of the call.  Uses the trace hook of
elasticdeform_b200.deform_grid (_TRACE), i.e. the very code path a NumPy caller takes.

    python scripts/e2e_timeline.py [slabs] [sigma] > gpurun_out/e2e_timeline.txt
"""
import sys, os, time, importlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import elasticdeform_b200 as edf
dg = importlib.import_module("elasticdeform_b200.deform_grid")

SLABS = int(sys.argv[1]) if len(sys.argv) > 1 else 8
SIGMA = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
dg._PIPELINE_SLABS = SLABS
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
N = 256
rng = np.random.default_rng(0)
Xp = torch.empty((N,) * 3, dtype=torch.float32).pin_memory()
Xp.copy_(torch.from_numpy(rng.random((N,) * 3, dtype=np.float32)))
Xn = Xp.numpy()
D = rng.standard_normal((3, 5, 5, 5)) * SIGMA
held = []
for name, fn in (("forward", lambda: edf.deform_grid(Xn, D, order=3, prefilter=False)),
                 ("gradient", lambda: edf.deform_grid_gradient(Xn, D, order=3, prefilter=False))):
    for i in range(4):
        dg._TRACE = [] if i == 3 else None
        t0 = time.perf_counter()
        held.append(fn())
        t1 = time.perf_counter()
        held = held[-2:]
    tr, dg._TRACE = dg._TRACE, None
    e0 = tr[0][1]
    print("%s, %d slabs, sigma %g: host wall %.3f ms" % (name, SLABS, SIGMA, (t1 - t0) * 1e3))
    for label, e in tr[1:]:
        print("  %8.3f ms  %s" % (e0.elapsed_time(e), label))
