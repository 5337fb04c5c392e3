#!/bin/bash
# staged forward A/B + sanitizer: bash scripts/gpu_quick2.sh <tag>
TAG=$1
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x -k "staged_window_forward" > gpurun_out/pytest_q_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_q_$TAG.log)
tail -3 gpurun_out/pytest_q_$TAG.log
EDF_STAGED_FWD=1 python scripts/ab_time.py 3,2,5 8,16 > gpurun_out/abq_$TAG.jsonl 2> gpurun_out/abq_$TAG.err
cut -c1-130 gpurun_out/abq_$TAG.jsonl; tail -2 gpurun_out/abq_$TAG.err
(timeout 600 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x -k "staged_window_forward and (nearest-7.0-shape1-3] or constant-100.0-shape6-3] or nearest-600.0-shape7-3])" > gpurun_out/memcheck_$TAG.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck_$TAG.log)
tail -3 gpurun_out/memcheck_$TAG.log
(timeout 600 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -x -k "staged_window_forward and (nearest-7.0-shape1-3] or constant-7.0-shape0-3])" > gpurun_out/racecheck_$TAG.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/racecheck_$TAG.log)
tail -3 gpurun_out/racecheck_$TAG.log
(timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:edf_swin3d_fwd' -s 6 -c 1 \
   -f -o gpurun_out/prof_swinfwd_$TAG env EDF_STAGED_FWD=1 python scripts/ab_time.py 3 8 > gpurun_out/ncu_swinfwd_$TAG.log 2>&1; echo "ncu-full rc=$?")
