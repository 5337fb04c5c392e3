#!/bin/bash
# GPU visit for the staged-window kernels: parity, A/B timing against the round-1 kernels, sanitizer, ncu.
# Usage (under gpurun): bash scripts/gpu_swin.sh <tag> [full-tests]
TAG=${1:-sw1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
(timeout 900 python -m pytest tests -m gpu -q -x -k "staged_window or window_gradient" > gpurun_out/pytest_swin_$TAG.log 2>&1; echo "pytest-swin rc=$?" >> gpurun_out/pytest_swin_$TAG.log)
tail -5 gpurun_out/pytest_swin_$TAG.log
python scripts/ab_time.py 3,2,5 8,16 >> gpurun_out/ab_$TAG.jsonl 2>> gpurun_out/ab_$TAG.err
EDF_SWIN_GRAD_MAXORDER=-1 python scripts/ab_time.py 3,2,5 8,16 >> gpurun_out/ab_${TAG}_old.jsonl 2>> gpurun_out/ab_$TAG.err
python scripts/ab_time.py 0,1 8,16 >> gpurun_out/ab_$TAG.jsonl 2>> gpurun_out/ab_$TAG.err
EDF_SWIN_GRAD_MAXORDER=-1 python scripts/ab_time.py 0,1 8,16 >> gpurun_out/ab_${TAG}_old.jsonl 2>> gpurun_out/ab_$TAG.err
echo "--- staged"; cut -c1-200 gpurun_out/ab_$TAG.jsonl; echo "--- round-1 kernels"; cut -c1-200 gpurun_out/ab_${TAG}_old.jsonl
(timeout 600 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x -k "staged_window_gradient and (nearest-7.0-shape0-3] or constant-100.0-shape4-3] or nearest-600.0-shape5-3])" > gpurun_out/memcheck_$TAG.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck_$TAG.log)
tail -4 gpurun_out/memcheck_$TAG.log
(timeout 600 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -x -k "staged_window_gradient and (nearest-7.0-shape0-3] or constant-100.0-shape4-3])" > gpurun_out/racecheck_$TAG.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/racecheck_$TAG.log)
tail -4 gpurun_out/racecheck_$TAG.log
(timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:edf_swin3d_grad' -s 6 -c 1 \
   -f -o gpurun_out/prof_swingrad_$TAG python scripts/ab_time.py 3 8 > gpurun_out/ncu_swin_$TAG.log 2>&1; echo "ncu-full rc=$?")
if [ "$2" == "full-tests" ]; then
  (timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log)
  tail -6 gpurun_out/pytest_gpu_$TAG.log
fi
ls -la gpurun_out | tail -6
