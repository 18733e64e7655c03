#!/bin/bash
# one GPU call: GPU tests + smoke at head, then A/B of the STFT's evict-first bulk stores (scratch/variants/lib_ef.so)
python -m pytest tests -q -m gpu -x > gpurun_out/r2_v10_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_v10_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/r2_v10_smoke.log 2>&1; tail -1 gpurun_out/r2_v10_smoke.log
cp audio-style-transfer_b200/libast_frontend.so /tmp/lib_orig.so
for v in base ef base ef; do
  if [ $v == base ]; then cp /tmp/lib_orig.so audio-style-transfer_b200/libast_frontend.so; else cp scratch/variants/lib_$v.so audio-style-transfer_b200/libast_frontend.so; fi
  echo -n "$v "; python scratch/prof_step.py --steps 100 --legs features,stats
done > gpurun_out/ef_ab.log 2>&1
cp /tmp/lib_orig.so audio-style-transfer_b200/libast_frontend.so
cat gpurun_out/ef_ab.log
