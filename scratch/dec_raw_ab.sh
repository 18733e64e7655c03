#!/bin/bash
for d in 32 0 32 0; do echo -n "AST_DEC_DEBUG=$d "; AST_DEC_DEBUG=$d python scratch/prof_step.py --steps 50 --legs features,stats --profile; done
