#!/bin/bash
# usage: scratch/ncu_kernel.sh <kernel regex> <tag> [legs]   - one ncu --set full capture (with source) of the first
# matching launch after warm-up, preceded by the plain run the profiling guide asks for
LEGS=${3:-features}
python scratch/prof_step.py --steps 5 --legs $LEGS --profile > gpurun_out/$2_step.json 2> gpurun_out/$2_step.err && \
ncu --set full --clock-control none --import-source on -k regex:$1 -s 3 -c 1 -f -o gpurun_out/$2 python scratch/prof_step.py --steps 1 --legs $LEGS > gpurun_out/$2_ncu.log 2>&1
cat gpurun_out/$2_step.json; tail -2 gpurun_out/$2_ncu.log
