#!/bin/bash
# usage: scripts/gpu_ncu.sh <tag> <kernel regex> <launch-skip> <count> <command ...>
# one `ncu --set full` capture (run on the GPU box); exports the raw and source pages as CSV into gpurun_out/
TAG=$1; KRE=$2; SKIP=$3; CNT=$4; shift 4
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG "$@" > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_src.csv 2>/dev/null
ls -la gpurun_out/prof_$TAG*
# keep the report only if it is small enough to travel back
SZ=$(stat -c %s gpurun_out/prof_$TAG.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 30000000 ]; then rm -f gpurun_out/prof_$TAG.ncu-rep; fi
