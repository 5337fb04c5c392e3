import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import elasticdeform_b200 as edf
from elasticdeform_b200 import _lib
np.set_printoptions(linewidth=200, precision=5, suppress=True)
S = (64, 64, 64)
rng = np.random.default_rng(1)
D = rng.standard_normal((3, 5, 5, 5)) * 0.5
for trial in range(3):
    G = np.zeros(S, np.float32)
    pts = [(8 + 13 * trial, 9 + 7 * trial, 10 + 5 * trial), (40, 41 + trial, 42)]
    for pnt in pts:
        G[pnt] = 1.0 + trial
    a = edf.deform_grid_gradient(G, D, order=3); ka = _lib.last_kernel()
    b = edf.deform_grid_gradient(G, D, order=3, _flags=_lib.EDF_FLAG_NO_WINDOW)
    d = a - b
    print("trial", trial, ka, "pts", pts, "maxdiff %.3e" % np.abs(d).max(), "sum a %.7f b %.7f" % (a.sum(), b.sum()), "nnz a", (a != 0).sum(), "nnz b", (b != 0).sum())
    bad = np.argwhere(np.abs(d) > 1e-7)
    for i in bad[:40]:
        print("   ", tuple(int(v) for v in i), "a %.7f b %.7f" % (a[tuple(i)], b[tuple(i)]))
# dense small-amplitude noise + one big voxel
G = (rng.standard_normal(S) * 1e-3).astype(np.float32); G[20, 21, 22] = 5.0
a = edf.deform_grid_gradient(G, D, order=3); b = edf.deform_grid_gradient(G, D, order=3, _flags=_lib.EDF_FLAG_NO_WINDOW)
print("big voxel in noise: maxdiff %.3e" % np.abs(a - b).max())
G = rng.standard_normal(S).astype(np.float32)
a = edf.deform_grid_gradient(G, D, order=3); b = edf.deform_grid_gradient(G, D, order=3, _flags=_lib.EDF_FLAG_NO_WINDOW)
d = np.abs(a - b); print("dense normal: maxdiff %.3e" % d.max(), "n>3e-5", (d > 3e-5).sum())
bad = np.argwhere(d > 3e-5)
for i in bad[:30]:
    print("   ", tuple(int(v) for v in i), "a %.7f b %.7f" % (a[tuple(i)], b[tuple(i)]))
