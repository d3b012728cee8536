#!/bin/bash
# A/B of (library, environment) configurations on the GPU box: tools/ab2.sh "<lib or ->[,VAR=val...]" ...
# e.g. tools/ab2.sh "-,CPN_ESTEP_PAIRS=0" "-" "build/variants/ep3.so"
for cfg in "$@"; do
  IFS=',' read -ra parts <<< "$cfg"
  ( L=${parts[0]}
    if [ "$L" != "-" ]; then export CPN_LIB=$PWD/$L; fi
    for kv in "${parts[@]:1}"; do export "$kv"; done
    python bench.py --steps 3 --warmup 2 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', round(d['value']), round(d['ms_per_step'],2), d['kernel_ms_per_step'], 'e2e', d['e2e']['ms_steps_rank0'], 'stats', d['fit_stats'], 'frac', d['roofline']['frac'])" )
done
