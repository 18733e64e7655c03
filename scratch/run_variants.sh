#!/bin/bash
# usage: scratch/run_variants.sh p0 p1 ...   (libraries built into scratch/variants/lib_<tag>.so)
# Swaps each library in turn into the package and prints the per-kernel times of a short bench run.
for v in "$@"; do
  cp scratch/variants/lib_$v.so audio-style-transfer_b200/libast_frontend.so
  for i in 1 2; do
    python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', {k:round(v['ms_per_step'],4) for k,v in d['roofline']['kernels'].items()}, round(d['ms_per_step'],4), round(d['istft']['ms_per_step'],4))"
  done
done
