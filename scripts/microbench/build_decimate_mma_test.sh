#!/bin/bash
# builds scripts/microbench/decimate_mma_test (tcgen05 decimator vs float64 FIR vs the FFMA2 kernel)
set -e
cd "$(dirname "$0")/../.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Wno-deprecated-gpu-targets $DM_FLAGS \
  -I ser_b200/csrc scripts/microbench/decimate_mma_test.cu ser_b200/csrc/decimate_mma.cu ser_b200/csrc/cqt_kernels.cu \
  ser_b200/csrc/cqt_tables.cpp ser_b200/csrc/filterbanks.cpp -o scripts/microbench/decimate_mma_test
