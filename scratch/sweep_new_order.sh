#!/bin/bash
for v in 16 24 32 48; do echo -n "AST_DEC_TAIL_CTAS=$v "; AST_DEC_TAIL_CTAS=$v python scratch/prof_step.py --steps 50 --legs features,stats; done
