#!/bin/bash
# A/B of the decimator kernels (AST_DECIMATOR = tf32 | f16; AST_DEC_DEBUG=16: f16 without the L2 prefetch): step times
for v in tf32 f16 f16; do
  d=0; k=$v; if [ $v == f16np ]; then d=16; k=f16; fi
  echo -n "AST_DECIMATOR=$v "; AST_DEC_DEBUG=$d AST_DECIMATOR=$k python scratch/prof_step.py --steps 50 --legs features,stats --profile
done
