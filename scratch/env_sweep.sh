#!/bin/bash
# usage: scratch/env_sweep.sh VAR v1 v2 ...   - per-kernel times of a short bench run for each value of an env switch
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$var=$v', {k:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items()}, round(d['ms_per_step'],4), round(d['istft']['ms_per_step'],4))"
done
