#!/bin/bash
# A/B timing of library variants on the GPU box: tools/ab.sh <regions> lib1.so lib2.so ...   ("-" = in-tree library)
R=$1; shift
for L in "$@"; do
  if [ "$L" = "-" ]; then unset CPN_LIB; else export CPN_LIB=$PWD/$L; fi
  python bench.py --regions $R --steps 2 --warmup 1 --cpu-sample 16 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$L', round(d['value']), round(d['ms_per_step'],2), d['kernel_ms_per_step'])"
done
