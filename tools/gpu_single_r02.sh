#!/bin/bash
# 1 GPU: new-length tests, the whole GPU suite, the default bench line and a launch list of the fused spectral operator
OUT=gpurun_out/r02e
mkdir -p $OUT
(time python -m pytest tests/test_gpu_lengths.py tests/test_z_programs_gpu.py tests/test_gpu_pair_conv.py -m gpu -q -s) > $OUT/lengths.log 2>&1; tail -12 $OUT/lengths.log
(time python -m pytest tests -m gpu -q --deselect tests/test_gpu_lengths.py --deselect tests/test_z_programs_gpu.py --deselect tests/test_gpu_pair_conv.py) > $OUT/pytest.log 2>&1; tail -6 $OUT/pytest.log
python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -2 $OUT/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_conv.csv python bench.py --kind conv --steps 2 --warmup 1 --no-secondary --no-parity --no-e2e > $OUT/ncu_conv.log 2>&1; tail -3 $OUT/ncu_conv.log
timeout 300 tools/kbench_real > $OUT/kbench_real.log 2>&1; grep -A8 "real2" $OUT/kbench_real.log | head -60
