#!/bin/bash
# A/B of the feature call's launch order (AST_FEATURE_ORDER = dcs: decimator, CQT, STFT | dsc: decimator, STFT, CQT)
for v in dcs dsc dcs dsc; do
  echo -n "AST_FEATURE_ORDER=$v "; AST_FEATURE_ORDER=$v python scratch/prof_step.py --steps 50 --legs features,stats
done
