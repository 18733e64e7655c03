#!/bin/bash
# A/B of the decimator's chain-stage dealing (AST_DEC_TAIL_CTAS: 0 = round-robin over the whole grid)
for v in 32 24 16 32; do
  echo -n "AST_DEC_TAIL_CTAS=$v "; AST_DEC_TAIL_CTAS=$v python scratch/prof_step.py --steps 50 --legs features,stats --profile
done
