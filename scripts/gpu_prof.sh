#!/bin/bash
# quick visit: bench + full ncu capture of the forward and gradient kernels
TAG=${1:-p}
mkdir -p gpurun_out
(timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?" >> gpurun_out/bench_$TAG.err)
python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print(d['value'], d['kernels'], d['e2e']['value'], d['parity'])"
(timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:edf_(fast_f32|lean3d)' -s 4 -c 2 \
   -f -o gpurun_out/prof_$TAG python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu-full rc=$?")
