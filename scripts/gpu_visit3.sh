#!/bin/bash
# GPU visit: e2e probe (PCIe rates, one-shot vs pipelined API call) + rows-per-CTA A/B of the staged-window kernels.
TAG=${1:-v3}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/smi_$TAG.txt 2>&1
lscpu | head -20 >> gpurun_out/smi_$TAG.txt
timeout 600 python scripts/e2e_probe.py > gpurun_out/e2e_probe_$TAG.jsonl 2> gpurun_out/e2e_probe_$TAG.err
cat gpurun_out/e2e_probe_$TAG.jsonl | cut -c1-220
for R in 0 16 8; do
  echo "--- EDF_SWIN_ROWS=$R"
  EDF_SWIN_ROWS=$R timeout 300 python scripts/ab_time.py 3,2,1 8 2>> gpurun_out/ab_$TAG.err | tee -a gpurun_out/ab_${TAG}_rows$R.jsonl | cut -c1-200
done
(EDF_SWIN_ROWS=16 timeout 600 python -m pytest tests -m gpu -q -x -k "staged_window" > gpurun_out/pytest_rows16_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_rows16_$TAG.log)
tail -3 gpurun_out/pytest_rows16_$TAG.log
(timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:edf_swin3d_fwd' -s 6 -c 1 \
   -f -o gpurun_out/prof_swinfwd_$TAG python scripts/ab_time.py 3 8 > gpurun_out/ncu_swinfwd_$TAG.log 2>&1; echo "ncu-full fwd rc=$?")
ls -la gpurun_out | tail -8
ncu -i gpurun_out/prof_swinfwd_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_swinfwd_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_swinfwd_$TAG.ncu-rep --page source --csv > gpurun_out/prof_swinfwd_${TAG}_src.csv 2>/dev/null
ls -la gpurun_out/prof_swinfwd_$TAG* ; S=$(stat -c %s gpurun_out/prof_swinfwd_$TAG.ncu-rep); if [ "$S" -gt 40000000 ]; then rm gpurun_out/prof_swinfwd_$TAG.ncu-rep; fi
