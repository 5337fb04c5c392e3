#!/bin/bash
# GPU visit: new prologue / y-contraction of the staged-window kernels -- parity, same-box A/B against the previous
# build (_variants/), configuration matrix.
TAG=${1:-v6}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "staged_window or default_kernel or headline or cfg or golden" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log)
tail -4 gpurun_out/pytest_$TAG.log
bash scripts/ab_variants.sh 3,2,1 8 2>> gpurun_out/ab_$TAG.err | tee gpurun_out/ab_$TAG.jsonl | cut -c1-175
bash scripts/ab_variants.sh 3 8 0 2>> gpurun_out/ab_$TAG.err | tee -a gpurun_out/ab_$TAG.jsonl | cut -c1-175
timeout 600 python scripts/bench_matrix.py > gpurun_out/matrix_$TAG.jsonl 2>> gpurun_out/ab_$TAG.err
cut -c1-230 gpurun_out/matrix_$TAG.jsonl
tail -3 gpurun_out/ab_$TAG.err
for V in "EDF_SWIN_ROWS=32" "EDF_SWIN_ROWS=32 EDF_SWIN_TAIL=5" "EDF_SWIN_ROWS=16 EDF_SWIN_TAIL=5" "EDF_SWIN_ROWS=16 EDF_SWIN_TAIL=10" "EDF_SWIN_ROWS=8"; do
  echo "--- $V"
  env $V timeout 300 python scripts/ab_time.py 3,2 8 2>> gpurun_out/ab_$TAG.err | sed "s/^{/{\"env\": \"$V\", /" | tee -a gpurun_out/ab_${TAG}_sched.jsonl | cut -c1-200
done
