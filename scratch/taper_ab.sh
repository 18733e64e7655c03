#!/bin/bash
# STFT tapered tail (AST_STFT_TAPER = "half[,quarter]" clips dealt as half / quarter rows at the end of the grid)
for v in ${TAPERS:-0 8 6,2 8,4 4,4 12,4 8,8 0,8 4,2}; do echo -n "AST_STFT_TAPER=$v "; AST_STFT_TAPER=$v python scratch/prof_step.py --steps 100 --legs features; done > gpurun_out/taper_ab3.log 2>&1
cat gpurun_out/taper_ab3.log

