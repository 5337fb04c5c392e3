import time, sys, os, importlib
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import elasticdeform_b200 as edf
dg = importlib.import_module("elasticdeform_b200.deform_grid")
S = (256,) * 3
rng = np.random.default_rng(0)
Xp = torch.empty(S, dtype=torch.float32).pin_memory(); Xp.copy_(torch.from_numpy(rng.random(S, dtype=np.float32)))
Xn = Xp.numpy(); D = rng.standard_normal((3, 5, 5, 5)) * 8
# instrument
marks = []
orig_pf = dg._pipelined_forward
def wrap(name, fn):
    def w(*a, **k):
        t = time.perf_counter(); r = fn(*a, **k); marks.append((name, (time.perf_counter() - t) * 1e3)); return r
    return w
dg._pinned_result = wrap("pinned_result", dg._pinned_result)
dg._launch = wrap("launch", dg._launch)
dg._prefilter_displacement = wrap("prefilter_disp", dg._prefilter_displacement)
dg._reach = wrap("reach", dg._reach)
dg._pipelined_forward = wrap("pipelined_forward", dg._pipelined_forward)
orig_empty = torch.empty
def timed_empty(*a, **k):
    t = time.perf_counter(); r = orig_empty(*a, **k); marks.append(("torch.empty%s" % ("(pin)" if k.get("pin_memory") else ("(dev)" if k.get("device") is not None else "")), (time.perf_counter() - t) * 1e3)); return r
torch.empty = timed_empty
keep = None
for it in range(5):
    marks.clear()
    torch.cuda.synchronize(); t = time.perf_counter()
    keep = edf.deform_grid(Xn, D, order=3, prefilter=False)
    dt = (time.perf_counter() - t) * 1e3
    agg = {}
    for n, v in marks: agg[n] = agg.get(n, 0) + v
    print("iter %d total %.2f ms:" % (it, dt), {k: round(v, 2) for k, v in agg.items()})
