#!/bin/bash
# ncu evidence of the training step with the fused inter data gradient: launch list of ONE step (+ DRAM bytes) and a
# full capture of the fused backward kernel on its three layers.
mkdir -p gpurun_out
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file gpurun_out/r4_ncu_step_launches.csv python bench.py --profile-step > gpurun_out/r4_step.log 2>&1
echo "launch list rc=$?"
timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:inter_bwd_fused_kernel -c 3 \
  -f -o gpurun_out/r4_bwd_fused python bench.py --profile-step > gpurun_out/r4_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/ | tail -8
