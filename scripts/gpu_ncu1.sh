#!/bin/bash
# one ncu --set full capture: bash scripts/gpu_ncu1.sh <tag> <kernel regex> [env assignments...] -- <script args>
TAG=$1; RX=$2; shift 2
mkdir -p gpurun_out
ENVS=()
while [ "$1" != "--" ] && [ -n "$1" ]; do ENVS+=("$1"); shift; done
shift
(timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$RX" -s 6 -c 1 \
   -f -o gpurun_out/prof_$TAG env "${ENVS[@]}" python scripts/ab_time.py "$@" > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu-full rc=$?")
ls -la gpurun_out | tail -4
