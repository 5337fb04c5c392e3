#!/usr/bin/env python
"""Chunk statistics of the pipelined forward kernel (library built with -DEDF_PIPE_STATS):
    python -m elasticdeform_b200.build --out _variants/libstats.so -DEDF_PIPE_STATS
    EDF_B200_LIB=_variants/libstats.so python scripts/pipe_stats.py [order] [sigma]
"""
import sys, os, importlib, ctypes, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from elasticdeform_b200 import _lib
dg = importlib.import_module("elasticdeform_b200.deform_grid")
order = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sigma = float(sys.argv[2]) if len(sys.argv) > 2 else 8.0
lib = _lib.load_library(); dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
X = torch.from_numpy(rng.random((256,) * 3, dtype=np.float32)).to(dev)
Y = torch.empty_like(X)
d_f = dg._prefilter_displacement(lib, rng.standard_normal((3, 5, 5, 5)) * sigma, dev)
pf, keep = dg._build_problem([X], [Y], d_f, None, [(0, 1, 2)], np.array([order]), np.array([4]), np.array([0.0]), None)
sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
out = (ctypes.c_uint64 * 16)()
lib.edf_debug_tile_profile.argtypes = [ctypes.POINTER(ctypes.c_uint64)]
lib.edf_debug_tile_profile(out)
_lib.check(lib.edf_deform_grid(ctypes.byref(pf), sp)); torch.cuda.synchronize()
lib.edf_debug_tile_profile(out)
_lib.check(lib.edf_deform_grid(ctypes.byref(pf), sp)); torch.cuda.synchronize()
lib.edf_debug_tile_profile(out)
ch = max(1, out[10])
print(json.dumps({"kernel": _lib.last_kernel(), "order": order, "sigma": sigma, "chunks": out[10], "rows_per_chunk": out[11] / ch,
                  "unstaged_chunks": out[9], "stage_rows_per_chunk": out[12] / ch, "rare_voxels": out[8], "voxels": X.numel(), "cta_Mcyc_max": out[13] / 1e6, "cta_Mcyc_mean": out[14] / 1e6 / 148,
                  "Mcyc": {"cons_wait_full": out[0] / 1e6, "cons_compute": out[1] / 1e6, "cons_tables_misc": out[2] / 1e6, "cons_rare": out[3] / 1e6, "cons_patch": out[7] / 1e6,
                           "prod_sample": out[4] / 1e6, "prod_wait_empty": out[5] / 1e6, "prod_issue": out[6] / 1e6}}))
