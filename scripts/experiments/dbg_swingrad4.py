import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import elasticdeform_b200 as edf
from elasticdeform_b200 import _lib
S = (96, 112, 128)
for order in (5, 3):
    for sigma in (7.0, 1.0):
        for pf in (False, True):
            rng = np.random.default_rng(4200 + order)
            G = rng.standard_normal(S).astype(np.float32)
            D = rng.standard_normal((3, 4, 5, 5)) * sigma
            a = edf.deform_grid_gradient(G, D, order=order, prefilter=pf); ka = _lib.last_kernel()
            f = edf.deform_grid_gradient(G, D, order=order, prefilter=pf, _flags=_lib.EDF_FLAG_FIXED_WINDOW)
            b = edf.deform_grid_gradient(G, D, order=order, prefilter=pf, _flags=_lib.EDF_FLAG_NO_WINDOW)
            b2 = edf.deform_grid_gradient(G, D, order=order, prefilter=pf, _flags=_lib.EDF_FLAG_NO_WINDOW)
            da, df, db = (a - b).astype(np.float64), (f - b).astype(np.float64), (b2 - b).astype(np.float64)
            print("order %d sigma %.0f prefilter %d: staged std %.2e max %.2e | fixed std %.2e max %.2e | direct rerun std %.2e max %.2e | max|b| %.1f"
                  % (order, sigma, pf, da.std(), np.abs(da).max(), df.std(), np.abs(df).max(), db.std(), np.abs(db).max(), np.abs(b).max()))
