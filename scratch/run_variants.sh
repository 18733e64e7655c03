#!/bin/bash
# usage: scratch/run_variants.sh [--legs features,istft] p0 p1 ...   (libraries built by scratch/build_variant.sh)
# Swaps each library in turn into the package and prints the CUDA-event times of scratch/prof_step.py (twice each).
LEGS=features
if [ "$1" == "--legs" ]; then LEGS=$2; shift 2; fi
cp audio-style-transfer_b200/libast_frontend.so /tmp/lib_orig.so
for v in "$@"; do
  cp scratch/variants/lib_$v.so audio-style-transfer_b200/libast_frontend.so
  for i in 1 2; do echo -n "$v "; python scratch/prof_step.py --steps 30 --legs $LEGS --profile; done
done
cp /tmp/lib_orig.so audio-style-transfer_b200/libast_frontend.so
