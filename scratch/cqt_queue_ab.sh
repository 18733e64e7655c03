#!/bin/bash
# A/B of the CQT projection's tile dealing (AST_CQT_STATIC=1: round-robin) in both launch orders
for o in dsc dcs; do for q in static queue static queue; do
  if [ $q == static ]; then export AST_CQT_STATIC=1; else unset AST_CQT_STATIC; fi
  echo -n "order $o cqt $q "; AST_FEATURE_ORDER=$o python scratch/prof_step.py --steps 50 --legs features,stats
done; done
