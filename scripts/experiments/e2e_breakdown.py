#!/usr/bin/env python
"""Where does the NumPy-path (e2e) time go?  Forward and gradient API calls timed separately, next to the raw
pinned H2D / D2H copy times of one 256^3 float32 volume on the same box (the PCIe floor of a call is one upload
overlapped with one download)."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import elasticdeform_b200 as edf

S = (256, 256, 256)
rng = np.random.default_rng(0)
Xp = torch.empty(S, dtype=torch.float32).pin_memory(); Xp.copy_(torch.from_numpy(rng.random(S, dtype=np.float32)))
Gp = torch.empty(S, dtype=torch.float32).pin_memory(); Gp.copy_(torch.from_numpy(rng.random(S, dtype=np.float32)))
D = rng.standard_normal((3, 5, 5, 5)) * 8.0
Xn, Gn = Xp.numpy(), Gp.numpy()
dev = torch.device("cuda", 0)
out = {}

def timeit(fn, n=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return round(float(np.median(ts)), 3), round(float(np.min(ts)), 3)

Xd = torch.empty(S, dtype=torch.float32, device=dev)
Yp = torch.empty(S, dtype=torch.float32).pin_memory()
out["h2d_67MB_ms"] = timeit(lambda: Xd.copy_(Xp, non_blocking=True))
out["d2h_67MB_ms"] = timeit(lambda: Yp.copy_(Xd, non_blocking=True))
s2 = torch.cuda.Stream()
def duplex():
    Xd.copy_(Xp, non_blocking=True)
    with torch.cuda.stream(s2):
        Yp.copy_(Xd, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)
out["duplex_67MB_each_ms"] = timeit(duplex)
for o in (3, 1):
    out["fwd_o%d_ms" % o] = timeit(lambda: edf.deform_grid(Xn, D, order=o, prefilter=False))
    out["grad_o%d_ms" % o] = timeit(lambda: edf.deform_grid_gradient(Gn, D, order=o, prefilter=False))
out["fwd_o3_default_prefilter_ms"] = timeit(lambda: edf.deform_grid(Xn, D, order=3))
out["grad_o3_default_prefilter_ms"] = timeit(lambda: edf.deform_grid_gradient(Gn, D, order=3))
Xt = Xd
out["fwd_o3_cuda_tensor_ms"] = timeit(lambda: edf.deform_grid(Xt, D, order=3, prefilter=False))
out["cpu_count"] = os.cpu_count()
print(json.dumps(out))
