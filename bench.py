#!/usr/bin/env python
"""bench.py -- headline benchmark of the elastic-deformation hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json `metric`): forward + gradient of a 256^3 float32 volume at
spline order 3 with a 5x5x5 displacement grid (sigma 8), mode 'constant',
prefilter=False (kernel-only headline).  One "step" = one forward gather +
the zero-fill of dX + one gradient scatter of one volume per GPU; volumes are
independent, so N GPUs process N volumes per step (weak scaling, no data-path
collective; NCCL only for the barrier / max-over-ranks).

Printed keys (one JSON line from rank 0):
  value         Mvoxels/s fwd+grad, data resident in HBM, C-ABI calls, CUDA-event timed
  e2e           same metric through the public Python API with pinned HOST arrays
                (H2D of X and dY, D2H of Y and dX inside the timed region)
  roofline      the dominant kernel's algorithmic bytes / its CUDA-event duration vs the
                measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own C loop (oracle/_ref; else the oracle port) timed on
                this box's host cores on a bounded sample of the same workload
`--impl reference` times only that CPU path and prints the same line with
"impl": "reference".
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = (256, 256, 256)
POINTS = (5, 5, 5)
SIGMA = 8.0
ORDER = 3
NVOX = SHAPE[0] * SHAPE[1] * SHAPE[2]
ALG_BYTES_FWD = 8 * NVOX          # read X once + write Y once (float32)        SURVEY 8d
ALG_BYTES_GRAD = 8 * NVOX         # read dY once + write dX once (zero-fill not credited)
METRIC = "Mvoxels/s fwd+grad (256^3 f32 order=3)"


def make_inputs(seed):
    rng = np.random.default_rng(seed)
    X = rng.random(SHAPE, dtype=np.float32)
    dY = rng.random(SHAPE, dtype=np.float32)
    D = rng.standard_normal((3,) + POINTS) * SIGMA
    return X, dY, D


# ----------------------------------------------------------------------------------------
# CPU reference arm (also the cpu_baseline leg of the GPU arm)
# ----------------------------------------------------------------------------------------
def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_reference_rate(X, dY, D, seconds_budget, planes=None, steps=1, max_threads=None, whole=True):
    """fwd+grad Mvoxels/s of the reference C loop on the host cores.

    Every host thread deforms its own slab of z-planes of the SAME volume -- the reference's own crop
    mechanism (output_offset, deform.c:438-446) computes exactly those output voxels with the
    displacement field and input of the full volume -- forward and gradient, through the reference's C
    entry points (_deform_grid.deform_grid / deform_grid_grad, i.e. oracle/_ref; the oracle port when
    the reference did not travel).  The C loop releases the GIL (deform.c:377-379), so the threads run
    in parallel.  ``whole=True``: the slabs tile the volume (planes = ceil(256 / threads)), so ONE step
    is the whole 256^3 workload; the timed region then also holds what the reference's Python layer
    does around the loop for a gradient: the zero-fill of the accumulators (deform_grid.py:243) and,
    because each thread scatters into its own full-size accumulator, their sum.  ``whole=False`` with
    ``planes``: a bounded sample of `threads x planes` z-planes (accumulators zeroed outside the timed
    region: in a slab sample the fill would dominate).
    """
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    kind = "reference" if O.ref_available() else "port"
    mod, pf = O._backend("ref" if kind == "reference" else "port")
    cores = max_threads or os.cpu_count() or 1
    cores = min(cores, SHAPE[0])
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
        cores = max(1, min(cores, int(avail * 0.5 // (4 * NVOX))))
    except (ValueError, OSError):
        pass
    if whole:
        planes = -(-SHAPE[0] // cores)
        starts = [k * planes for k in range(cores) if k * planes < SHAPE[0]]
    else:
        planes = planes or 4
        nslab = SHAPE[0] // planes
        cores = min(cores, nslab)
        starts = [(k * (nslab // cores)) * planes for k in range(cores)]
    njobs = len(starts)
    d_f = O._prefilter_displacement(D, pf)
    axis = [(0, 1, 2)]
    order, mode, cval = np.array([ORDER]), np.array([4]), np.array([0.0])
    dXs = [np.zeros(SHAPE, np.float32) for _ in range(njobs)]
    for a in dXs:
        a.fill(0.0)                                        # touch the pages
    outs = [np.zeros((min(planes, SHAPE[0] - z0),) + SHAPE[1:], np.float32) for z0 in starts]

    def job(k):
        z0 = starts[k]
        n = outs[k].shape[0]
        off = np.array([z0, 0, 0], dtype=np.int64) if z0 > 0 else None
        mod.deform_grid([X], d_f, off, [outs[k]], axis, order, mode, cval, None)
        if whole:
            dXs[k].fill(0.0)
        mod.deform_grid_grad([dXs[k]], d_f, off, [dY[z0:z0 + n]], axis, order, mode, cval, None)
        return outs[k].size

    def reduce_job(k):                                      # sum of the per-thread accumulators, by z-range
        z0 = starts[k]
        n = outs[k].shape[0]
        acc = dXs[0][z0:z0 + n]
        for j in range(1, njobs):
            acc += dXs[j][z0:z0 + n]
        return 0

    times = []
    with ThreadPoolExecutor(njobs) as ex:
        list(ex.map(job, range(min(2, njobs))))            # warm (page-in, lazy builds)
        t_all0 = time.perf_counter()
        for _ in range(steps):
            t0 = time.perf_counter()
            vox = sum(ex.map(job, range(njobs)))
            if whole and njobs > 1:
                list(ex.map(reduce_job, range(njobs)))
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_all0 > seconds_budget:
                break
    t = float(np.mean(times))
    if whole:
        sample = ("the whole 256^3 f32 order-3 volume per step: %d host threads x %d z-planes through the reference's "
                  "crop offset, fwd + dX zero-fill + grad + sum of the per-thread accumulators = %d voxels per step"
                  % (njobs, planes, vox))
    else:
        sample = ("%d host threads x (%d z-planes of the 256^3 f32 order-3 volume, fwd+grad via the "
                  "reference's crop offset) = %d voxels per step" % (njobs, planes, vox))
    return vox / t / 1e6, njobs, kind, sample, t, len(times)


def best_cpu_reference_rate(X, dY, D, seconds_budget, steps, whole=True):
    """The reference with all the host threads it can use: one software thread per hardware thread
    and one per two (SMT siblings share the FP units; which is faster depends on the host)."""
    ncpu = os.cpu_count() or 1
    best = None
    for thr in sorted({ncpu, max(1, ncpu // 2)}, reverse=True):
        if whole:
            r = cpu_reference_rate(X, dY, D, seconds_budget / 2, steps=steps, max_threads=thr, whole=True)
        else:
            planes = max(1, min(4, SHAPE[0] // thr))
            r = cpu_reference_rate(X, dY, D, seconds_budget / 2, planes=planes, steps=steps, max_threads=thr, whole=False)
        if best is None or r[0] > best[0]:
            best = r
    return best


def single_thread_rate(X, dY, D):
    """One host thread, 2 z-planes of the same volume (~0.5 s): the reference as a user calls it."""
    v, _, _, _, _, _ = cpu_reference_rate(X, dY, D, 20.0, planes=2, steps=1, max_threads=1, whole=False)
    return v


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    X, dY, D = make_inputs(0)
    # one step = the whole volume (about 3-4 s on 16-32 threads); warm-up + steps bounded to a few minutes in total
    v, cores, kind, sample, t, nsteps = best_cpu_reference_rate(X, dY, D, seconds_budget=150.0,
                                                                steps=max(1, min(args.steps, 3)), whole=True)
    v1 = single_thread_rate(X, dY, D)
    line = {
        "metric": METRIC, "value": round(v, 4), "unit": "Mvoxels/s", "impl": "reference",
        "n_gpus": args.gpus, "steps": nsteps, "warmup": 1, "ms_per_step": round(t * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "256^3 float32 volume, 5x5x5 grid sigma 8, order 3, mode constant, "
                               "prefilter=False; step = forward gather + dX zero-fill + gradient scatter",
                   "host": sample, "cpu_model": cpu_model(), "host_threads_available": os.cpu_count()},
        "cpu_baseline": {"value": round(v, 4), "unit": "Mvoxels/s", "cores": cores, "kind": kind,
                         "sample": sample, "single_thread_value": round(v1, 4), "cpu_model": cpu_model()},
        "e2e": {"value": round(v, 4), "unit": "Mvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------
# clocks sampler (NVML; falls back to nvidia-smi)
# ----------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
               0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.active = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self._stop_evt.is_set():
            try:
                mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                if self.active:
                    self.samples.append(mhz)
                    for bit, name in self.REASONS.items():
                        if r & bit and name != "gpu_idle":
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import importlib
    import elasticdeform_b200 as edf
    from elasticdeform_b200 import _lib
    dg = importlib.import_module("elasticdeform_b200.deform_grid")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"           # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load_library()

    # ---- resident data: NSETS buffer sets rotate so that no step re-reads L2-resident inputs
    NSETS = 4
    X_h, dY_h, D = make_inputs(rank)
    sets = []
    for s in range(NSETS):
        X = torch.from_numpy(np.roll(X_h, s, axis=0)).to(dev)
        dY = torch.from_numpy(np.roll(dY_h, s, axis=1)).to(dev)
        Y = torch.empty_like(X)
        dX = torch.empty_like(X)
        sets.append((X, dY, Y, dX))
    d_f = dg._prefilter_displacement(lib, D, dev)
    axis = [(0, 1, 2)]
    order, mode, cval = np.array([ORDER]), np.array([4]), np.array([0.0])
    probs = []
    for (X, dY, Y, dX) in sets:
        pf, kf = dg._build_problem([X], [Y], d_f, None, axis, order, mode, cval, None)
        pg, kg = dg._build_problem([dX], [dY], d_f, None, axis, order, mode, cval, None)
        probs.append((pf, pg, kf, kg))
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)

    def step(i, evs=None):
        X, dY, Y, dX = sets[i % NSETS]
        pf, pg, _, _ = probs[i % NSETS]
        if evs is not None:
            evs[0].record(stream)
        _lib.check(lib.edf_deform_grid(ctypes.byref(pf), sp))
        if evs is not None:
            evs[1].record(stream)
        dX.zero_()
        if evs is not None:
            evs[2].record(stream)
        _lib.check(lib.edf_deform_grid_grad(ctypes.byref(pg), sp))
        if evs is not None:
            evs[3].record(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    sampler.start()

    _lib.check(lib.edf_deform_grid(ctypes.byref(probs[0][0]), sp))
    fwd_kernel = _lib.last_kernel()
    _lib.check(lib.edf_deform_grid_grad(ctypes.byref(probs[0][1]), sp))
    grad_kernel = _lib.last_kernel()
    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()

    # ---- timed region: exactly K steps, CUDA events on the launching stream
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.launch_count()
    sampler.active = True
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step(i, evs[i])
    e1.record(stream)
    barrier()
    sampler.active = False
    launches = _lib.launch_count() - launches0 + args.steps      # + the zero-fill kernels (torch)
    ms_total = e0.elapsed_time(e1)
    t_fwd = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    t_zero = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    t_grad = float(np.mean([e[2].elapsed_time(e[3]) for e in evs]))
    tmax = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    ms_step = ms_total / args.steps
    value = NVOX * world / (ms_step * 1e-3) / 1e6

    # ---- parity spot check on the very arrays that were timed (a slab, vs the oracle)
    parity = None
    if rank == 0:
        try:
            from oracle import oracle as O
            impl = "ref" if O.ref_available() else "port"
            crop = (slice(128, 130), slice(0, 256), slice(0, 256))
            X, dY, Y, dX = sets[0]
            step(0)
            torch.cuda.synchronize(dev)
            yref = O.deform_grid(X.cpu().numpy(), D, order=ORDER, prefilter=False, crop=crop, impl=impl)
            parity = {"fwd_max_abs_err_slab": float(np.abs(Y[128:130].cpu().numpy() - yref).max()), "tol": 1e-5}
        except Exception as e:                                   # pragma: no cover
            parity = {"error": repr(e)}

    # ---- e2e: public API, pinned host arrays in and out, every step
    Xp = torch.empty(SHAPE, dtype=torch.float32).pin_memory()
    Gp = torch.empty(SHAPE, dtype=torch.float32).pin_memory()
    Xp.copy_(torch.from_numpy(X_h))
    Gp.copy_(torch.from_numpy(dY_h))
    Xn, Gn = Xp.numpy(), Gp.numpy()
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step():
        y = edf.deform_grid(Xn, D, order=ORDER, prefilter=False)
        dx = edf.deform_grid_gradient(Gn, D, order=ORDER, prefilter=False)
        return y, dx

    for _ in range(3):                          # results held across the next call, as in the timed loop, so
        y, dx = e2e_step()                      # that the pinned result pool reaches its steady-state size here
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        y, dx = e2e_step()
    torch.cuda.synchronize(dev)
    t_e2e = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())
    e2e_value = NVOX * world / t_e2e / 1e6

    # default-API variant (prefilter=True) for information
    for _ in range(2):
        y = edf.deform_grid(Xn, D, order=ORDER)
        dx = edf.deform_grid_gradient(Gn, D, order=ORDER)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(3):
        y = edf.deform_grid(Xn, D, order=ORDER)
        dx = edf.deform_grid_gradient(Gn, D, order=ORDER)
    torch.cuda.synchronize(dev)
    t_e2e_pf = (time.perf_counter() - t0) / 3

    # pageable NumPy inputs (what a user of the reference passes: no pinning on the caller's side)
    Xq, Gq = np.array(X_h), np.array(dY_h)
    for _ in range(2):
        y = edf.deform_grid(Xq, D, order=ORDER, prefilter=False)
        dx = edf.deform_grid_gradient(Gq, D, order=ORDER, prefilter=False)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(3):
        y = edf.deform_grid(Xq, D, order=ORDER, prefilter=False)
        dx = edf.deform_grid_gradient(Gq, D, order=ORDER, prefilter=False)
    torch.cuda.synchronize(dev)
    t_e2e_pageable = (time.perf_counter() - t0) / 3

    # torch wrapper on CUDA tensors: forward + backward, nothing crosses PCIe (the reference's wrapper copies the
    # tensor to the host, runs the C loop and copies the result back, in both directions: torch.py:13-16, :25-29)
    import elasticdeform_b200.torch as etorch
    Xt = sets[0][0].clone().requires_grad_(True)
    Gt = sets[0][1]
    Dt = torch.from_numpy(D)

    def torch_step():
        Xt.grad = None
        Yt = etorch.deform_grid(Xt, Dt, order=ORDER, prefilter=False)
        Yt.backward(Gt)

    for _ in range(3):
        torch_step()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(10):
        torch_step()
    torch.cuda.synchronize(dev)
    t_torch = (time.perf_counter() - t0) / 10

    sampler.stop()
    clocks = sampler.summary()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        if t_grad >= t_fwd:
            dom, t_dom, alg = "gradient scatter (" + grad_kernel + ")", t_grad, ALG_BYTES_GRAD
        else:
            dom, t_dom, alg = "forward gather (" + fwd_kernel + ")", t_fwd, ALG_BYTES_FWD
        achieved = alg / (t_dom * 1e-3) / 1e9
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get("grad" if t_grad >= t_fwd else "fwd")
        except Exception:
            pass
        ach_f = ALG_BYTES_FWD / (t_fwd * 1e-3) / 1e9
        ach_g = ALG_BYTES_GRAD / (t_grad * 1e-3) / 1e9
        cpu = None                                               # timed at N=1 only (rank 0); null in the N>1 lines
        if world == 1:
            try:
                v, cores, kind, sample, t, _ = best_cpu_reference_rate(X_h, dY_h, D, seconds_budget=30.0, steps=1, whole=False)
                cpu = {"value": round(v, 4), "unit": "Mvoxels/s", "cores": cores, "kind": kind, "sample": sample,
                       "single_thread_value": round(single_thread_rate(X_h, dY_h, D), 4), "cpu_model": cpu_model()}
            except Exception as e:                               # pragma: no cover
                cpu = {"error": repr(e)}
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "Mvoxels/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "256^3 float32 volume per GPU, 5x5x5 grid sigma 8, order 3, mode constant, "
                                   "prefilter=False; step = forward gather + dX zero-fill + gradient scatter",
                       "l2": "inputs larger than L2: %d rotating buffer sets (%.2f GB) vs 126 MB L2"
                             % (NSETS, NSETS * 4 * 4 * NVOX / 1e9),
                       "sharding": "one independent volume per GPU per step, no data-path collective"},
            "e2e": {"value": round(e2e_value, 1), "unit": "Mvoxels/s",
                    "h2d_bytes_per_step": 2 * 4 * NVOX, "d2h_bytes_per_step": 2 * 4 * NVOX,
                    "ms_per_step": round(t_e2e * 1e3, 3),
                    "api": "elasticdeform_b200.deform_grid + deform_grid_gradient on pinned NumPy arrays",
                    "default_prefilter_ms_per_step": round(t_e2e_pf * 1e3, 3),
                    "pageable_numpy_ms_per_step": round(t_e2e_pageable * 1e3, 3),
                    "torch_cuda_ms_per_step": round(t_torch * 1e3, 3),
                    "torch_cuda_value": round(NVOX / t_torch / 1e6, 1),
                    "torch_cuda_api": "elasticdeform_b200.torch.deform_grid on a CUDA tensor + backward (no host copies; "
                                      "the reference's wrapper round-trips through .cpu().numpy(), i.e. runs at cpu_baseline speed)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak,
                         "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg, "kernel_ms": round(t_dom, 4)},
            "kernels": {"fwd_ms": round(t_fwd, 4), "zero_fill_ms": round(t_zero, 4), "grad_ms": round(t_grad, 4),
                        "fwd_GBps": round(ach_f, 1), "fwd_frac": round(ach_f / peak, 4),
                        "grad_GBps": round(ach_g, 1), "grad_frac": round(ach_g / peak, 4),
                        "fwd_kernel": fwd_kernel, "grad_kernel": grad_kernel},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "parity": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ----------------------------------------------------------------------------------------
# Other BASELINE.json configs (`--config cfg3|cfg4|cfg5`): same line format, their own metric names
# ----------------------------------------------------------------------------------------
def cfg4_affine(n=128, crop0=32, crop1=96, deg=15.0, zoom=1.2):
    """rotation about axis 0 by 15 degrees and zoom 1.2, centred on the crop centre, as a (3, 4) affine."""
    th = np.deg2rad(deg)
    R = np.array([[1, 0, 0], [0, np.cos(th), -np.sin(th)], [0, np.sin(th), np.cos(th)]]) * zoom
    c = np.full(3, 0.5 * (crop0 + crop1 - 1))
    return np.concatenate([R, (c - R @ c)[:, None]], axis=1)


def run_config_arm(args):
    import torch
    import torch.distributed as dist
    import elasticdeform_b200 as edf
    from elasticdeform_b200 import _lib, batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream(dev)
    rng = np.random.default_rng(100 + rank)
    cfg = args.config
    NS = 3                                                      # rotating input sets (inputs >> L2 over a rotation)

    if cfg == "cfg3":
        # 256^3 float32 image (order 3) + int32 label (order 0) sharing one displacement; one pair per GPU
        D = rng.standard_normal((3, 5, 5, 5)) * SIGMA
        sets = [(torch.from_numpy(rng.random(SHAPE, dtype=np.float32)).to(dev),
                 torch.from_numpy(rng.integers(0, 5, SHAPE, dtype=np.int32)).to(dev)) for _ in range(NS)]
        vox_step = NVOX                                          # output voxel positions (two arrays each)
        alg = 16 * NVOX

        def step(i):
            return edf.deform_grid(list(sets[i % NS]), D, order=[3, 0], prefilter=False)
        host_in = [t.cpu().pin_memory().numpy() for t in sets[0]]

        def e2e_step():
            return edf.deform_grid(host_in, D, order=[3, 0], prefilter=False)
        metric = "Mvoxels/s forward (cfg3: 256^3 f32 order 3 + int32 order 0 pair)"
        workload = "256^3 float32 image + int32 label per GPU, orders [3, 0], one 5^3 displacement (sigma 8), prefilter=False"
        scaling, h2d, d2h = "weak", 8 * NVOX, 8 * NVOX
    elif cfg == "cfg4":
        # 64 volumes of 128^3 float32, each with its own displacement and 3-D affine, crop 64^3; the batch is cut
        # into contiguous blocks of volumes, one block per GPU (batch.shard_range): strong scaling over a fixed batch
        B, n = 64, 128
        b0, b1 = batch.shard_range(B, rank, world)
        crop = (slice(32, 96),) * 3
        A = cfg4_affine()
        grng = np.random.default_rng(4)
        Ds_all = [grng.standard_normal((3, 5, 5, 5)) * 4.0 for _ in range(B)]
        Ds = Ds_all[b0:b1]
        sets = [[torch.from_numpy(rng.random((n,) * 3, dtype=np.float32)).to(dev) for _ in range(b1 - b0)] for _ in range(NS)]
        vox_step = B * 64 ** 3
        alg = B * (4 * 64 ** 3 + 4 * 64 ** 3)                    # touched input ~ the cropped region + output

        def step(i):
            return batch.deform_grid_batch(sets[i % NS], Ds, order=3, crop=crop, prefilter=False, affines=[A] * len(Ds))
        host_in = [t.cpu().pin_memory().numpy() for t in sets[0]]

        def e2e_step():
            return batch.deform_grid_batch(host_in, Ds, order=3, crop=crop, prefilter=False, affines=[A] * len(Ds))
        metric = "Mvoxels/s forward, output voxels (cfg4: batch 64 x 128^3 f32 -> crop 64^3, affine + 5^3 grid, order 3)"
        workload = ("batch of 64 volumes 128^3 float32, per-volume displacement (sigma 4) + 3-D affine (15 deg, zoom 1.2), "
                    "crop 64^3, order 3, prefilter=False; volumes sharded over the GPUs in contiguous blocks")
        scaling, h2d, d2h = "strong", (b1 - b0) * 4 * n ** 3, (b1 - b0) * 4 * 64 ** 3
    else:
        # cfg5: one array of 32 channels x 128^3, axis=(1, 2, 3), order 1, forward + gradient; channels sharded
        C, n = 32, 128
        c0, c1 = batch.shard_range(C, rank, world)
        D = np.random.default_rng(5).standard_normal((3, 5, 5, 5)) * 4.0
        sets = [(torch.from_numpy(rng.random((c1 - c0, n, n, n), dtype=np.float32)).to(dev),
                 torch.from_numpy(rng.random((c1 - c0, n, n, n), dtype=np.float32)).to(dev)) for _ in range(NS)]
        vox_step = C * n ** 3
        alg = C * n ** 3 * 16

        def step(i):
            X, G = sets[i % NS]
            y = edf.deform_grid(X, D, order=1, axis=(1, 2, 3), prefilter=False)
            dx = edf.deform_grid_gradient(G, D, order=1, axis=(1, 2, 3), prefilter=False)
            return y, dx
        host_in = [t.cpu().pin_memory().numpy() for t in sets[0]]

        def e2e_step():
            y = edf.deform_grid(host_in[0], D, order=1, axis=(1, 2, 3), prefilter=False)
            dx = edf.deform_grid_gradient(host_in[1], D, order=1, axis=(1, 2, 3), prefilter=False)
            return y, dx
        metric = "Mvoxels/s fwd+grad, channel voxels (cfg5: 32 x 128^3 f32, axis=(1,2,3), order 1)"
        workload = ("one array of 32 channels x 128^3 float32 sharing one 5^3 displacement (sigma 4), axis=(1,2,3), order 1, "
                    "forward + gradient, prefilter=False; channels sharded over the GPUs in contiguous blocks")
        scaling, h2d, d2h = "strong", 2 * (c1 - c0) * 4 * n ** 3, 2 * (c1 - c0) * 4 * n ** 3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    sampler.start()
    for i in range(max(args.warmup, 3)):
        step(i)
    kernel = _lib.last_kernel()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.launch_count()
    sampler.active = True
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    sampler.active = False
    launches = _lib.launch_count() - launches0
    tmax = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / args.steps
    value = vox_step / (ms_step * 1e-3) / 1e6

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(3):
        keep = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        keep = e2e_step()
    torch.cuda.synchronize(dev)
    te = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())
    sampler.stop()
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0)) * world
        achieved = alg / (ms_step * 1e-3) / 1e9
        line = {
            "metric": metric, "value": round(value, 1), "unit": "Mvoxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "name": cfg,
                       "l2": "%d rotating input sets per GPU" % NS,
                       "api": "public Python API on CUDA tensors (value) / pinned NumPy arrays (e2e)"},
            "e2e": {"value": round(vox_step / t_e2e / 1e6, 1), "unit": "Mvoxels/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": round(t_e2e * 1e3, 3)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "whole step (" + kernel + ")", "achieved": round(achieved, 1), "peak": peak,
                         "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": None,
                         "algorithmic_bytes_per_launch": int(alg)},
            "cpu_baseline": None,
            "clocks": sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="headline", choices=["headline", "cfg3", "cfg4", "cfg5"],
                    help="headline = BASELINE.json's metric (256^3 f32 order 3 fwd+grad); cfg3-5 = its other GPU configs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.config != "headline":
        return run_config_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
