#!/bin/bash
# 1 GPU, about a minute: variant tables of the c2c kernels (512^3 fp64, 256^3 fp32) and of the real-data kernels.
# Build first, here or in the CPU container (nvcc cross-compiles):
#   for t in kbench_c2c kbench_real; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/$t tools/$t.cu; done
OUT=gpurun_out/variants
mkdir -p $OUT
timeout 200 tools/kbench_c2c > $OUT/kbench_c2c.log 2>&1; echo "kbench_c2c rc=$?"
timeout 200 tools/kbench_real > $OUT/kbench_real.log 2>&1; echo "kbench_real rc=$?"
cat $OUT/kbench_c2c.log $OUT/kbench_real.log
