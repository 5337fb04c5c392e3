#!/bin/bash
# GPU visit: the slab pipeline with per-slab displacement bounds -- parity, timeline, probe.
TAG=${1:-v4}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x -k "slab_pipelined or host or torch_wrapper" > gpurun_out/pytest_pipe_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pipe_$TAG.log)
tail -15 gpurun_out/pytest_pipe_$TAG.log
python scripts/e2e_timeline.py 8 8 > gpurun_out/e2e_timeline8_$TAG.txt 2>&1
python scripts/e2e_timeline.py 16 8 > gpurun_out/e2e_timeline16_$TAG.txt 2>&1
cat gpurun_out/e2e_timeline8_$TAG.txt
timeout 600 python scripts/e2e_probe.py > gpurun_out/e2e_probe_$TAG.jsonl 2> gpurun_out/e2e_probe_$TAG.err
grep api gpurun_out/e2e_probe_$TAG.jsonl | cut -c1-200; tail -3 gpurun_out/e2e_probe_$TAG.err
