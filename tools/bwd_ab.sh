#!/bin/bash
# Fused inter data gradient inside ONE box: its parity test, then the training step over knob settings
# (tools/bwd_ab.sh "EPN_FUSED_BWD=0" "EPN_FUSED_BWD=1" "EPN_FB_SPS=2" ...).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_inter_data_gradient or full_size or many_slabs" > gpurun_out/bwd_test.log 2>&1
echo "test rc=$?" >> gpurun_out/bwd_test.log
tail -4 gpurun_out/bwd_test.log
for kv in "$@"; do
  env $kv timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 8 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$kv', round(d['ms_per_step'], 2), d['kernel_ms_per_step'], d['clocks']['sm_mhz'])"
done 2>&1 | tee gpurun_out/bwd_ab.log
