#!/bin/bash
# diagnostic: decimator kernel time with parts of the pipeline disabled (AST_DEC_DEBUG bit mask)
for d in ${@:-0 1 2 4 3 7}; do
  AST_DEC_DEBUG=$d python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('debug=$d', 'dec', round(d['roofline']['kernels']['decimate2_tc_kernel']['ms_per_step'],4), 'step', round(d['ms_per_step'],4))"
done
