#!/usr/bin/env python
"""e2e fwd+grad time per step of the public API on pinned NumPy arrays for a given EDF_PIPELINE_SLABS (A/B runs)."""
import sys, os, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import elasticdeform_b200 as edf
rng = np.random.default_rng(0)
Xp = torch.empty((256,) * 3, dtype=torch.float32).pin_memory(); Gp = torch.empty((256,) * 3, dtype=torch.float32).pin_memory()
Xp.copy_(torch.from_numpy(rng.random((256,) * 3, dtype=np.float32))); Gp.copy_(torch.from_numpy(rng.random((256,) * 3, dtype=np.float32)))
Xn, Gn = Xp.numpy(), Gp.numpy()
D = rng.standard_normal((3, 5, 5, 5)) * 8
res = {}
for name, kw in (("prefilter_false", dict(prefilter=False)), ("default", dict())):
    for _ in range(3):
        y = edf.deform_grid(Xn, D, order=3, **kw); dx = edf.deform_grid_gradient(Gn, D, order=3, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        y = edf.deform_grid(Xn, D, order=3, **kw)
        t1 = time.perf_counter()
        dx = edf.deform_grid_gradient(Gn, D, order=3, **kw)
        torch.cuda.synchronize()
        ts.append((t1 - t0, time.perf_counter() - t1))
    res[name] = {"fwd_ms": round(1e3 * float(np.median([a for a, b in ts])), 3), "grad_ms": round(1e3 * float(np.median([b for a, b in ts])), 3)}
print(json.dumps({"slabs": os.environ.get("EDF_PIPELINE_SLABS", "8"), **res}))
