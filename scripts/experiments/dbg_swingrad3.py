import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import elasticdeform_b200 as edf
from elasticdeform_b200 import _lib
np.set_printoptions(linewidth=220, precision=2, suppress=True)
S = (32, 128, 256) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split(","))
rng = np.random.default_rng(1)
D = rng.standard_normal((3, 5, 5, 5)) * 0.5
G = rng.standard_normal(S).astype(np.float32)
a = edf.deform_grid_gradient(G, D, order=3, prefilter=False); ka = _lib.last_kernel()
b = edf.deform_grid_gradient(G, D, order=3, prefilter=False, _flags=_lib.EDF_FLAG_NO_WINDOW)
d = a - b
print(S, ka, "maxdiff %.3e" % np.abs(d).max(), "n>1e-5", (np.abs(d) > 1e-5).sum(), "sum d %.4e" % d.astype(np.float64).sum())
bad = np.argwhere(np.abs(d) > 1e-5)
if len(bad):
    print("z", bad[:, 0].min(), bad[:, 0].max(), "y", bad[:, 1].min(), bad[:, 1].max(), "x", bad[:, 2].min(), bad[:, 2].max())
    import collections
    print("by y%4", sorted(collections.Counter(bad[:, 1] % 4).items()), "by y//4 (first 20)", sorted(collections.Counter(bad[:, 1] // 4).items())[:20])
    print("by x%32", sorted(collections.Counter(bad[:, 2] % 32).items()))
    i = bad[len(bad) // 2]
    z, y, x = [int(v) for v in i]
    print("around", (z, y, x), "(units 1e-6)")
    print((d[max(0, z - 3):z + 4, max(0, y - 3):y + 4, max(0, x - 3):x + 4] * 1e6))
