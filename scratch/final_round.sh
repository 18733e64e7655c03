#!/bin/bash
# last evidence call of the round: GPU tests + smoke at head, then scratch/profile_round2.sh (bench both arms, launch list, ncu full)
python -m pytest tests -q -m gpu -x > gpurun_out/r2_v10_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_v10_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/r2_v10_smoke.log 2>&1; tail -1 gpurun_out/r2_v10_smoke.log
bash scratch/profile_round2.sh > gpurun_out/r2_profile_round.log 2>&1
tail -c 600 gpurun_out/r2_bench.json
