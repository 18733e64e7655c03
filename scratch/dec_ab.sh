#!/bin/bash
# A/B of the decimator kernels (AST_DECIMATOR = tf32 | f16): step times, then the GPU tests on the default
for v in tf32 f16 tf32 f16; do
  echo -n "AST_DECIMATOR=$v "; AST_DECIMATOR=$v python scratch/prof_step.py --steps 50 --legs features,stats --profile
done
