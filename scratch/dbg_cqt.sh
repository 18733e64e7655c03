#!/bin/bash
# diagnostic: CQT kernel time with parts of the pipeline disabled (AST_CQT_DEBUG bit mask: 1 no epilogue stores,
# 2 no producer loads (register path), 4 no MMAs, 8 no L2 prefetch)
for d in ${@:-0 1 4 5 8}; do
  echo -n "debug=$d "; AST_CQT_DEBUG=$d python scratch/prof_step.py --steps 20 --legs features --profile
done
