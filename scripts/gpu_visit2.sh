#!/bin/bash
# e2e breakdown + ncu full capture of the window gradient kernel
TAG=${1:-v2}
mkdir -p gpurun_out
(timeout 300 python scripts/experiments/e2e_breakdown.py > gpurun_out/e2e_$TAG.json 2> gpurun_out/e2e_$TAG.err; echo "e2e rc=$?")
cat gpurun_out/e2e_$TAG.json; tail -3 gpurun_out/e2e_$TAG.err
(timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:gradwin' -s 6 -c 1 \
   -f -o gpurun_out/prof_gradwin_$TAG python scripts/ab_time.py 3 8 > gpurun_out/ncu_gradwin_$TAG.log 2>&1; echo "ncu-full rc=$?")
(timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:edf_lean3d_fwd' -s 6 -c 1 \
   -f -o gpurun_out/prof_leanfwd_$TAG env EDF_NO_SWIN=1 python scripts/ab_time.py 3 8 > gpurun_out/ncu_leanfwd_$TAG.log 2>&1; echo "ncu-full rc=$?")
ls -la gpurun_out | tail -5
