#!/usr/bin/env python
"""Time the 256^3 float32 forward / gradient kernels of the library named by EDF_B200_LIB (kernel A/B runs).

    EDF_B200_LIB=_variants/libA.so python scripts/ab_time.py [orders] [sigmas]

Prints one JSON line per (order, sigma): device times (CUDA events, median of 20 after 5 warm-ups, rotating over 3
buffer sets so that consecutive launches do not hit a warm L2) and a checksum of the results for a quick
cross-variant sanity check (parity proper is tests/test_parity_gpu.py).
"""
import sys, os, importlib, ctypes, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from elasticdeform_b200 import _lib
dg = importlib.import_module("elasticdeform_b200.deform_grid")
orders = [int(o) for o in (sys.argv[1] if len(sys.argv) > 1 else "3").split(",")]
sigmas = [float(o) for o in (sys.argv[2] if len(sys.argv) > 2 else "8").split(",")]
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 4      # 4 = constant
lib = _lib.load_library(); dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
NSET = 3
Xs = [torch.from_numpy(rng.random((256,) * 3, dtype=np.float32)).to(dev) for _ in range(NSET)]
Ys = [torch.empty_like(Xs[0]) for _ in range(NSET)]
dXs = [torch.zeros_like(Xs[0]) for _ in range(NSET)]
Dn = rng.standard_normal((3, 5, 5, 5))
sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
ax = [(0, 1, 2)]
for sigma in sigmas:
    d_f = dg._prefilter_displacement(lib, Dn * sigma, dev)
    for order in orders:
        args = (np.array([order]), np.array([mode]), np.array([0.0]), None)
        pf = [dg._build_problem([Xs[i]], [Ys[i]], d_f, None, ax, *args) for i in range(NSET)]
        pg = [dg._build_problem([dXs[i]], [Xs[(i + 1) % NSET]], d_f, None, ax, *args) for i in range(NSET)]
        res = {}
        for name, probs, fn in (("fwd", pf, lib.edf_deform_grid), ("grad", pg, lib.edf_deform_grid_grad)):
            ts = []
            for it in range(25):
                i = it % NSET
                if name == "grad":
                    dXs[i].zero_()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(fn(ctypes.byref(probs[i][0]), sp))
                e1.record(); torch.cuda.synchronize()
                if it >= 5:
                    ts.append(e0.elapsed_time(e1))
            res[name] = round(float(np.median(ts)), 4)
            res[name + "_kernel"] = _lib.last_kernel()
        res["sum_Y"] = float(Ys[0].double().sum().item())
        res["sum_dX"] = float(dXs[0].double().sum().item())
        res["absum_dX"] = float(dXs[0].double().abs().sum().item())
        print(json.dumps({"lib": os.path.basename(_lib.library_path()), "order": order, "sigma": sigma, **res}), flush=True)
