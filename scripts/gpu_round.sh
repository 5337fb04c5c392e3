#!/bin/bash
# One GPU-box visit: tests, bench, ncu launch list, ncu full capture of the two hot kernels.
# Usage (under gpurun): bash scripts/gpu_round.sh <tag> [skip-tests]
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  (timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log)
  tail -15 gpurun_out/pytest_gpu_$TAG.log
fi
(timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?" >> gpurun_out/bench_$TAG.err)
tail -2 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
# every launch of the same command with its device time (cold-cache, serialised)
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
   --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 5 --warmup 3 > gpurun_out/bench_under_ncu_$TAG.log 2>&1; echo "ncu-list rc=$?")
# full capture of the forward gather and the gradient scatter (one launch each)
(timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:edf_(swin3d|lean3d|fast_f32)' -s 4 -c 2 \
   -f -o gpurun_out/prof_$TAG python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu-full rc=$?")
# two full captures exceed the 64 MiB return limit: export the pages that are read afterwards, drop the report
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_src.csv 2>/dev/null
rm -f gpurun_out/prof_$TAG.ncu-rep
(timeout 600 python scripts/bench_matrix.py > gpurun_out/matrix_$TAG.jsonl 2> gpurun_out/matrix_$TAG.err; echo "matrix rc=$?")
cut -c1-200 gpurun_out/matrix_$TAG.jsonl
ls -la gpurun_out | tail -12
