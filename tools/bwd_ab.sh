#!/bin/bash
# A/B of the fused inter data gradient inside ONE box: test first, then the training step with the knob off / on.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_inter_data_gradient" > gpurun_out/bwd_test.log 2>&1
echo "test rc=$?" >> gpurun_out/bwd_test.log
tail -5 gpurun_out/bwd_test.log
for v in 0 1 0 1; do
  EPN_FUSED_BWD=$v timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 8 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('EPN_FUSED_BWD=$v', round(d['ms_per_step'], 2), d['kernel_ms_per_step'], d['clocks']['sm_mhz'])"
done 2>&1 | tee gpurun_out/bwd_ab.log
