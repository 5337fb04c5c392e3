import time, sys, os, importlib
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import elasticdeform_b200 as edf
dg = importlib.import_module("elasticdeform_b200.deform_grid")
S = (256,) * 3
rng = np.random.default_rng(0)
Xp = torch.empty(S, dtype=torch.float32).pin_memory(); Xp.copy_(torch.from_numpy(rng.random(S, dtype=np.float32)))
Xn = Xp.numpy(); D = rng.standard_normal((3, 5, 5, 5)) * 8
print("is_pinned(from_numpy):", torch.from_numpy(Xn).is_pinned())
def T(f, n=5):
    f(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3
print("pinned alloc 67MB: %.3f ms" % T(lambda: torch.empty(S, dtype=torch.float32, pin_memory=True)))
d = torch.empty(S, dtype=torch.float32, device="cuda")
print("H2D 67MB: %.3f ms" % T(lambda: d.copy_(torch.from_numpy(Xn), non_blocking=True)))
h = torch.empty(S, dtype=torch.float32, pin_memory=True)
print("D2H 67MB: %.3f ms" % T(lambda: h.copy_(d, non_blocking=True)))
print("fwd pipelined: %.3f ms" % T(lambda: edf.deform_grid(Xn, D, order=3, prefilter=False)))
print("grad pipelined: %.3f ms" % T(lambda: edf.deform_grid_gradient(Xn, D, order=3, prefilter=False)))
dg._PIPELINE_MIN_BYTES = 1 << 60
print("fwd one-shot: %.3f ms" % T(lambda: edf.deform_grid(Xn, D, order=3, prefilter=False)))
print("grad one-shot: %.3f ms" % T(lambda: edf.deform_grid_gradient(Xn, D, order=3, prefilter=False)))
Xd = torch.from_numpy(Xn).cuda()
print("fwd device-resident API: %.3f ms" % T(lambda: edf.deform_grid(Xd, D, order=3, prefilter=False)))
import cProfile, pstats
dg._PIPELINE_MIN_BYTES = 16 << 20
pr = cProfile.Profile(); pr.enable()
for _ in range(3): edf.deform_grid(Xn, D, order=3, prefilter=False)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
