#!/usr/bin/env python
"""Phase-cycle breakdown of the staged-window kernels (library built with -DEDF_TILE_PROFILE):
    EDF_B200_LIB=_variants/libprof.so python scripts/tile_profile.py [order] [sigma] [fwd|grad]"""
import sys, os, importlib, ctypes, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from elasticdeform_b200 import _lib
dg = importlib.import_module("elasticdeform_b200.deform_grid")
order = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sigma = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
what = sys.argv[3] if len(sys.argv) > 3 else "fwd"
lib = _lib.load_library(); dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
X = torch.from_numpy(rng.random((256,) * 3, dtype=np.float32)).to(dev)
Y = torch.empty_like(X); dX = torch.zeros_like(X)
D = rng.standard_normal((3, 5, 5, 5)) * sigma
d_f = dg._prefilter_displacement(lib, D, dev)
sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
args = (np.array([order]), np.array([4]), np.array([0.0]), None)
if what == "fwd":
    pr = dg._build_problem([X], [Y], d_f, None, [(0, 1, 2)], *args); fn = lib.edf_deform_grid
else:
    pr = dg._build_problem([dX], [X], d_f, None, [(0, 1, 2)], *args); fn = lib.edf_deform_grid_grad
out = (ctypes.c_uint64 * 16)()
lib.edf_debug_tile_profile.argtypes = [ctypes.POINTER(ctypes.c_uint64)]
for _ in range(3):
    _lib.check(fn(ctypes.byref(pr[0]), sp))
torch.cuda.synchronize()
lib.edf_debug_tile_profile(out)
N = 5
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N):
    _lib.check(fn(ctypes.byref(pr[0]), sp))
e1.record(); torch.cuda.synchronize()
lib.edf_debug_tile_profile(out)
v = [int(x) for x in out]
warps = max(v[15], 1)
tot = sum(v[:8]) or 1
names = ["prologue", "coords+box", "barrier", "stage+wait+patch", "gather/scatter", "cst/slow/flush", "p6", "p7"]
print(json.dumps({"kernel": _lib.last_kernel(), "order": order, "sigma": sigma, "ms": round(e0.elapsed_time(e1) / N, 4),
                  "cycles_per_warp": round(tot / warps), **{n: round(100.0 * c / tot, 1) for n, c in zip(names, v[:8]) if c}}))
