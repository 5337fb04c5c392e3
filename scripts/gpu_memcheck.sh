#!/bin/bash
# GPU visit: compute-sanitizer over the staged-window kernels (tile schedule, prologue) and the slab pipeline.
TAG=${1:-mc}
mkdir -p gpurun_out
(timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x \
   -k "default_kernel_choice or staged_window_forward_crop_affine or slab_pipelined or (staged_window_gradient and (nearest-7.0-shape0-3] or constant-100.0-shape4-3] or nearest-600.0-shape5-3])) or (staged_window_forward_against and mirror)" \
   > gpurun_out/memcheck_$TAG.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck_$TAG.log)
grep -E "passed|failed|ERROR SUMMARY|memcheck rc|Invalid|error" gpurun_out/memcheck_$TAG.log | head -12
(timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -x \
   -k "(staged_window_gradient and (nearest-7.0-shape0-3] or constant-100.0-shape4-3])) or (staged_window_forward_against and mirror and 3)" \
   > gpurun_out/racecheck_$TAG.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/racecheck_$TAG.log)
grep -E "passed|failed|RACECHECK SUMMARY|racecheck rc|hazard" gpurun_out/racecheck_$TAG.log | head -8
