#!/bin/bash
# Last GPU visit of the round: the complete GPU suite on the final build, smoke(), then memcheck on a subset.
TAG=${1:-fin}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log)
tail -6 gpurun_out/pytest_gpu_$TAG.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log); tail -2 gpurun_out/smoke_$TAG.log
python scripts/ab_time.py 3 8 2>/dev/null | cut -c1-200
(timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x \
   -k "staged_window_forward_crop_affine or (staged_window_gradient and (nearest-7.0-shape0-3] or constant-100.0-shape4-3])) or (staged_window_forward_against and mirror and 3)" \
   > gpurun_out/memcheck_$TAG.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck_$TAG.log)
grep -E "passed|failed|ERROR SUMMARY|memcheck rc|Invalid" gpurun_out/memcheck_$TAG.log | head -8
