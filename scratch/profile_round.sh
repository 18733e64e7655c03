#!/bin/bash
# Tests, default bench (both arms), ncu launch list, ncu --set full of the hot kernels.  One ncu pass per GPU call: run the
# blocks below as four separate calls (each ncu pass only after the same command has run plain).
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r1_v3_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'stft_kernel|decimate2_tc|cqt_tc' -s 24 -c 8 -f -o gpurun_out/r1_v3_features $CMD > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'istft_kernel|resample_2to1' -s 3 -c 1 -f -o gpurun_out/r1_v3_istft $CMD > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out | tail -8
