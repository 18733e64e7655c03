#!/bin/bash
# usage: scratch/build_variant.sh <tag> [-DNAME=VALUE ...]   -> scratch/variants/lib_<tag>.so (same flags as build.py)
tag=$1; shift
mkdir -p scratch/variants
C=audio-style-transfer_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" -Xptxas -v \
  -o scratch/variants/lib_$tag.so $C/plan.cu $C/stft.cu $C/decimate.cu $C/decimate_tc.cu $C/cqt.cu $C/cqt_tc.cu $C/istft.cu \
  $C/layout_stats.cu $C/resample.cu $C/metrics.cu $C/synth.cu $C/api.cu 2>&1 | grep -A2 "stft_kernelILi0\|istft_kernel" | grep -E "registers|spill"
