import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import elasticdeform_b200 as edf
from elasticdeform_b200 import _lib
def run(order, mode, sigma, shape, absG=False, seed=4200):
    rng = np.random.default_rng(seed + order)
    G = rng.standard_normal(shape).astype(np.float32)
    if absG: G = np.abs(G)
    D = rng.standard_normal((3, 4, 5, 5)) * sigma
    a = edf.deform_grid_gradient(G, D, order=order, mode=mode); ka = _lib.last_kernel()
    b = edf.deform_grid_gradient(G, D, order=order, mode=mode, _flags=_lib.EDF_FLAG_NO_WINDOW)
    f = edf.deform_grid_gradient(G, D, order=order, mode=mode, _flags=_lib.EDF_FLAG_FIXED_WINDOW)
    tol = 1e-5 * max(1.0, (order + 1) ** 3 / 64.0) * max(1.0, float(np.abs(b).max()))
    bad = np.argwhere(np.abs(a - b) > tol)
    print("order", order, mode, sigma, shape, "absG", absG, ka, "tol %.2e" % tol, "nbad", len(bad), "maxdiff %.3e" % np.abs(a - b).max(),
          "fixedwin maxdiff %.3e" % np.abs(f - b).max(), "sum a %.6f b %.6f" % (a.astype(np.float64).sum(), b.astype(np.float64).sum()))
    if len(bad):
        print("  z range", bad[:, 0].min(), bad[:, 0].max(), "y range", bad[:, 1].min(), bad[:, 1].max(), "x range", bad[:, 2].min(), bad[:, 2].max())
        for i in bad[:12]:
            print("   ", tuple(i), "a", a[tuple(i)], "b", b[tuple(i)])
        import collections
        print("  by z:", sorted(collections.Counter(bad[:, 0]).items())[:20])
        print("  by x%4:", sorted(collections.Counter(bad[:, 2] % 4).items()))
S = (96, 112, 128)
import sys
if len(sys.argv) > 1 and sys.argv[1] == "big":
    run(3, "constant", 8.0, (256, 256, 256))
    run(3, "constant", 2.0, (256, 256, 256))
    run(3, "constant", 0.5, (256, 256, 256))
else:
    run(2, "constant", 7.0, S)
    run(3, "constant", 2.0, S)
    run(3, "nearest", 7.0, S)
