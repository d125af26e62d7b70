#!/bin/bash
OUT=gpurun_out/single_quick
mkdir -p $OUT
( time timeout 600 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 200 python bench.py --steps 10 --warmup 3 --kind r2r --no-e2e --no-cpu-baseline > $OUT/bench_r2r_f64_512.log 2>&1; echo "bench r2r rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --kind r2c --no-e2e --no-cpu-baseline > $OUT/bench_r2c_f64_512.log 2>&1; echo "bench r2c rc=$?"
grep -h '"metric"' $OUT/bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'], '|', round(d['value'],1), 'GFlop/s', round(d['ms_per_step'],4), 'ms/step err', d['max_roundtrip_error'], 'roof', d['roofline'] and round(d['roofline']['frac'],3), [(s['kernel'], round(s['GB/s'])) for s in d['stages']])
"
