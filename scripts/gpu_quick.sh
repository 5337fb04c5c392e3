#!/bin/bash
# quick A/B: bash scripts/gpu_quick.sh <tag> <orders> <sigmas> [ENV=VAL ...]
TAG=$1; ORD=$2; SIG=$3; shift 3
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x -k "staged_window or window_gradient" > gpurun_out/pytest_q_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_q_$TAG.log)
tail -3 gpurun_out/pytest_q_$TAG.log
env "$@" python scripts/ab_time.py $ORD $SIG > gpurun_out/abq_$TAG.jsonl 2> gpurun_out/abq_$TAG.err
cut -c1-210 gpurun_out/abq_$TAG.jsonl
