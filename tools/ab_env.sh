#!/bin/bash
# A/B of an environment knob: ab_env.sh VAR v1 v2 ...
V=$1; shift
for x in "$@"; do
  export $V=$x
  python bench.py --steps 3 --warmup 3 --cpu-sample 16 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$V=$x', round(d['value']), round(d['ms_per_step'],2), d['kernel_ms_per_step'], d['e2e']['ms_steps_rank0'])"
done
