#!/bin/bash
# Round-2 evidence run (ONE gpurun call, one GPU): default bench (both arms), ncu launch list, ncu --set full of every hot
# kernel (feature call, iSTFT, statistics mode).  Every ncu pass follows a plain run of the same command (B200_PROFILING.md).
set -x
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_ref.json 2> gpurun_out/r2_ref.err
CMD="python scratch/prof_step.py --steps 1 --warmup 1 --legs features,istft,stats"
$CMD > gpurun_out/r2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'stft_kernel|decimate2_tc|cqt_tc|istft_kernel|stats_finalize' -c 20 -f -o gpurun_out/r2_full $CMD > gpurun_out/r2_ncu2.log 2>&1
ls -la gpurun_out | tail -6
