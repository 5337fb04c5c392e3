#!/usr/bin/env python
"""Run the 256^3 float32 forward (and gradient) of one spline order a few times (for ncu)."""
import sys, os, importlib, ctypes
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from elasticdeform_b200 import _lib
dg = importlib.import_module("elasticdeform_b200.deform_grid")
order = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
lib = _lib.load_library(); dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
X = torch.from_numpy(rng.random((256,) * 3, dtype=np.float32)).to(dev)
Y = torch.empty_like(X); dX = torch.zeros_like(X)
D = rng.standard_normal((3, 5, 5, 5)) * 8.0
d_f = dg._prefilter_displacement(lib, D, dev)
ax = [(0, 1, 2)]
pf, k1 = dg._build_problem([X], [Y], d_f, None, ax, np.array([order]), np.array([4]), np.array([0.0]), None)
pg, k2 = dg._build_problem([dX], [Y], d_f, None, ax, np.array([order]), np.array([4]), np.array([0.0]), None)
sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for _ in range(n):
    _lib.check(lib.edf_deform_grid(ctypes.byref(pf), sp))
    _lib.check(lib.edf_deform_grid_grad(ctypes.byref(pg), sp))
torch.cuda.synchronize()
print("done", _lib.last_kernel())
