#!/bin/bash
# GPU visit: graded-tail tile schedule of the staged-window kernels -- parity + A/B.
TAG=${1:-v5}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "staged_window or default_kernel or headline or cfg" > gpurun_out/pytest_tail_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tail_$TAG.log)
tail -4 gpurun_out/pytest_tail_$TAG.log
for T in 0 -1 8 3; do
  echo "--- EDF_SWIN_TAIL=$T"
  EDF_SWIN_TAIL=$T timeout 300 python scripts/ab_time.py 3,2,1 8 2>> gpurun_out/ab_$TAG.err | tee -a gpurun_out/ab_${TAG}_tail$T.jsonl | cut -c1-170
done
echo "--- EDF_SWIN_TAIL=-1 EDF_SWIN_ROWS=16"
EDF_SWIN_ROWS=16 timeout 300 python scripts/ab_time.py 3 8 2>> gpurun_out/ab_$TAG.err | tee -a gpurun_out/ab_${TAG}_tail_rows16.jsonl | cut -c1-170
echo "--- sigma 16 / mode nearest"
EDF_SWIN_TAIL=0 timeout 300 python scripts/ab_time.py 3 4,16 0 2>> gpurun_out/ab_$TAG.err | tee -a gpurun_out/ab_${TAG}_tail0_b.jsonl | cut -c1-170
timeout 300 python scripts/ab_time.py 3 4,16 0 2>> gpurun_out/ab_$TAG.err | tee -a gpurun_out/ab_${TAG}_taild_b.jsonl | cut -c1-170
tail -3 gpurun_out/ab_$TAG.err
