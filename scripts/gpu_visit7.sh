#!/bin/bash
# GPU visit: forward gather with the independent work moved behind the bulk-copy issue -- parity, same-box A/B, matrix.
TAG=${1:-v7}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "staged_window or default_kernel or headline or cfg or batch or edge" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log)
tail -4 gpurun_out/pytest_$TAG.log
bash scripts/ab_variants.sh 3,2,5 8 2>> gpurun_out/ab_$TAG.err | tee gpurun_out/ab_$TAG.jsonl | cut -c1-175
bash scripts/ab_variants.sh 3 8,16 0 2>> gpurun_out/ab_$TAG.err | tee -a gpurun_out/ab_$TAG.jsonl | cut -c1-175
timeout 600 python scripts/bench_matrix.py > gpurun_out/matrix_$TAG.jsonl 2>> gpurun_out/ab_$TAG.err
cut -c1-200 gpurun_out/matrix_$TAG.jsonl
tail -3 gpurun_out/ab_$TAG.err
