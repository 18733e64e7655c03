#!/bin/bash
for v in dcs dsc dcs dsc; do
  if [ $v == dcs ]; then export AST_STATS_ORDER_DCS=1; else unset AST_STATS_ORDER_DCS; fi
  echo -n "stats order $v "; python scratch/prof_step.py --steps 50 --legs features,stats
done
