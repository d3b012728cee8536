#!/bin/bash
# A/B timing of library variants on the GPU box: tools/ab.sh <regions> lib1.so lib2.so ...   ("-" = in-tree library)
R=$1; shift
for L in "$@"; do
  if [ "$L" = "-" ]; then unset CPN_LIB; else export CPN_LIB=$PWD/$L; fi
  python bench.py --regions $R --steps 3 --warmup 2 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$L', round(d['value']), round(d['ms_per_step'],2), d['kernel_ms_per_step'], 'e2e', d['e2e']['ms_steps_rank0'], 'conv', d['fit_stats']['converged_regions'], 'iters', d['fit_stats']['em_iters_per_region'])"
done
