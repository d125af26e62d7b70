#!/bin/bash
OUT=gpurun_out/debug
mkdir -p $OUT tests/emul/_build
g++ -std=c++17 -O1 -I include tests/cpp/example_b200.cpp -o $OUT/example_b200 -L heffte_b200/lib -lheffte_b200 -Wl,-rpath,$PWD/heffte_b200/lib -lpthread
HEFFTE_B200_TRACE=1 timeout 120 $OUT/example_b200 > $OUT/example_plain.log 2>&1; echo "example rc=$?"
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 $OUT/example_b200 > $OUT/example_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -m3 -B2 -A25 "Invalid\|Error" $OUT/example_memcheck.log | head -80
( time timeout 1700 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
