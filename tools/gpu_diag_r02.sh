#!/bin/bash
# 1 GPU: the sub-communicator plans on thread-ranks, the reference's reshape test repeated, the real-data kernel variants
OUT=gpurun_out/r02f
mkdir -p $OUT
(timeout 400 python -m pytest tests/test_z_programs_gpu.py -m gpu -q -k subcomm) > $OUT/subcomm.log 2>&1; echo "subcomm rc=$?"; tail -3 $OUT/subcomm.log
for i in 1 2 3 4 5 6 7 8; do (cd integration/_build; SHIM_NP=4 timeout 300 ./test_reshape3d > ../../$OUT/reshape3d_np4_$i.log 2>&1; echo -n "np4 run $i rc=$? "); done; echo
timeout 300 tools/kbench_real > $OUT/kbench_real.log 2>&1; grep -A7 "real2\|contig kind" $OUT/kbench_real.log | cut -c1-100
(time python -m pytest tests/test_gpu_fft1d.py tests/test_gpu_fft3d.py tests/test_y_fullsize_gpu.py -m gpu -q -x) > $OUT/pytest_fft.log 2>&1; tail -4 $OUT/pytest_fft.log
