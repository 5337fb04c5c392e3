#!/bin/bash
# GPU visit: visit 7's programme + the prefilter inside the slab pipeline.
TAG=${1:-v8}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "slab_pipelined or staged_window or default_kernel or headline or cfg or batch or edge" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log)
tail -12 gpurun_out/pytest_$TAG.log
bash scripts/ab_variants.sh 3,2,5 8 2>> gpurun_out/ab_$TAG.err | tee gpurun_out/ab_$TAG.jsonl | cut -c1-175
bash scripts/ab_variants.sh 3 8,16 0 2>> gpurun_out/ab_$TAG.err | tee -a gpurun_out/ab_$TAG.jsonl | cut -c1-175
python - > gpurun_out/e2e_pf_$TAG.txt 2>&1 <<'P'
import sys, time, importlib, numpy as np, torch
sys.path.insert(0, '.')
import elasticdeform_b200 as edf
dg = importlib.import_module("elasticdeform_b200.deform_grid")
rng = np.random.default_rng(0)
Xp = torch.empty((256,) * 3, dtype=torch.float32).pin_memory(); Xp.copy_(torch.from_numpy(rng.random((256,) * 3, dtype=np.float32)))
Xn = Xp.numpy(); D = rng.standard_normal((3, 5, 5, 5)) * 8
held = {}
def step():
    held['y'] = edf.deform_grid(Xn, D, order=3); held['g'] = edf.deform_grid_gradient(Xn, D, order=3)
def timed(label):
    for _ in range(3): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(6): step()
    torch.cuda.synchronize(); print(label, "fwd+grad, prefilter=True: %.3f ms per step" % ((time.perf_counter() - t0) / 6 * 1e3), flush=True)
timed("pipelined")
for name, fn in (("forward", lambda: edf.deform_grid(Xn, D, order=3)), ("gradient", lambda: edf.deform_grid_gradient(Xn, D, order=3))):
    dg._TRACE = []
    held[name] = fn()
    tr, dg._TRACE = dg._TRACE, None
    print(name, "timeline (ms):", "; ".join("%.2f %s" % (tr[0][1].elapsed_time(e), l) for l, e in tr[1:] if 'upload' not in l or l.endswith('7 done')))
keep = dg._PIPELINE_MIN_BYTES; dg._PIPELINE_MIN_BYTES = 1 << 60
timed("one-shot ")
P
cat gpurun_out/e2e_pf_$TAG.txt | cut -c1-1500
timeout 600 python scripts/bench_matrix.py > gpurun_out/matrix_$TAG.jsonl 2>> gpurun_out/ab_$TAG.err
cut -c1-200 gpurun_out/matrix_$TAG.jsonl
tail -3 gpurun_out/ab_$TAG.err
