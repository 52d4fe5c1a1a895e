#!/bin/bash
# A/B a tuning knob inside ONE gpurun box: tools/knob_ab.sh ENV_NAME v1 v2 ...   (prints step ms, GEMM class ms, SM MHz)
name=$1; shift
for v in "$@"; do
  env $name=$v python bench.py --no-extras --no-cpu-baseline --steps 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$name=$v', round(d['ms_per_step'],2), d['kernel_ms_per_step'], d['clocks']['sm_mhz'])"
done
