import time, sys, os, importlib, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from elasticdeform_b200 import _lib
dg = importlib.import_module("elasticdeform_b200.deform_grid")
lib = _lib.load_library(); dev = torch.device("cuda", 0)
S = (256,) * 3; h = 32
rng = np.random.default_rng(0)
Xp = torch.empty(S, dtype=torch.float32).pin_memory(); Xp.copy_(torch.from_numpy(rng.random(S, dtype=np.float32)))
Yp = torch.empty(S, dtype=torch.float32).pin_memory()
D = rng.standard_normal((3, 5, 5, 5)) * 8
d_f = dg._prefilter_displacement(lib, D, dev)
Xd = torch.empty(S, dtype=torch.float32, device=dev); Yd = torch.empty(S, dtype=torch.float32, device=dev)
cur = torch.cuda.current_stream(dev); s_up = torch.cuda.Stream(dev); s_down = torch.cuda.Stream(dev)
ax = [(0, 1, 2)]; od = np.array([3]); md = np.array([4]); cv = np.array([0.0])

def run(upload_slabs=True, kernels=True, downloads=True, reach=35, one_kernel=False):
    s_up.wait_stream(cur)
    evs = []
    with torch.cuda.stream(s_up):
        if upload_slabs:
            for j in range(8):
                Xd[j*h:(j+1)*h].copy_(Xp[j*h:(j+1)*h], non_blocking=True)
                e = torch.cuda.Event(); e.record(s_up); evs.append(e)
        else:
            Xd.copy_(Xp, non_blocking=True)
            e = torch.cuda.Event(); e.record(s_up); evs = [e] * 8
    if kernels:
        if one_kernel:
            cur.wait_event(evs[7])
            dg._launch(lib, 0, [Xd], [Yd], d_f, None, ax, od, md, cv, None)
            if downloads:
                e = torch.cuda.Event(); e.record(cur); s_down.wait_event(e)
                with torch.cuda.stream(s_down): Yp.copy_(Yd, non_blocking=True)
        else:
            for k in range(8):
                a, b = k*h, (k+1)*h
                need = min(255, b - 1 + reach)
                cur.wait_event(evs[need // h])
                dg._launch(lib, 0, [Xd], [Yd[a:b]], d_f, np.array([a, 0, 0], dtype='int64'), ax, od, md, cv, None)
                if downloads:
                    e = torch.cuda.Event(); e.record(cur); s_down.wait_event(e)
                    with torch.cuda.stream(s_down): Yp[a:b].copy_(Yd[a:b], non_blocking=True)
    cur.wait_stream(s_down); cur.wait_stream(s_up); cur.synchronize()

def T(f, n=6):
    f(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3

print("upload whole            %.2f ms" % T(lambda: run(False, False, False)))
print("upload 8 slabs          %.2f ms" % T(lambda: run(True, False, False)))
print("upload slabs + kernels  %.2f ms" % T(lambda: run(True, True, False)))
print("whole + 1 kernel        %.2f ms" % T(lambda: run(False, True, False, one_kernel=True)))
print("whole + 1 kernel + down %.2f ms" % T(lambda: run(False, True, True, one_kernel=True)))
print("full pipeline           %.2f ms" % T(lambda: run(True, True, True)))
print("full pipeline reach 0   %.2f ms" % T(lambda: run(True, True, True, reach=0)))
