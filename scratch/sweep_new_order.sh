#!/bin/bash
for v in 24 32 48; do echo -n "AST_DEC_TAIL_CTAS=$v "; AST_DEC_TAIL_CTAS=$v python scratch/prof_step.py --steps 50 --legs features,stats; done
for v in 2 3 4 6; do echo -n "AST_STFT_ITERS=$v "; AST_STFT_ITERS=$v python scratch/prof_step.py --steps 50 --legs features; done
