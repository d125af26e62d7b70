#!/bin/bash
# 1-GPU check of the final tree: GPU parity suite, smoke, headline bench line, real-kernel variants, ncu of the DCT stages
OUT=gpurun_out/single_final
mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 400 python bench.py --steps 20 --warmup 3 > $OUT/bench_c2c_f64_512.log 2>&1; echo "bench rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --kind r2r --no-e2e --no-cpu-baseline > $OUT/bench_r2r_f64_512.log 2>&1; echo "bench r2r rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --kind r2c --no-e2e --no-cpu-baseline > $OUT/bench_r2c_f64_512.log 2>&1; echo "bench r2c rc=$?"
timeout 200 tools/kbench_real > $OUT/kbench_real.log 2>&1; echo "kbench rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fft_ -c 3 -o $OUT/ncu_r2r_f64_512 -f \
    python bench.py --steps 1 --warmup 3 --kind r2r --no-e2e --no-cpu-baseline > $OUT/ncu_r2r.log 2>&1; echo "ncu r2r rc=$?"
grep -h '"metric"' $OUT/bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'], '|', round(d['value'],1), 'GFlop/s', round(d['ms_per_step'],4), 'ms/step err', d['max_roundtrip_error'], 'roof', d['roofline'] and round(d['roofline']['frac'],3), [(s['kernel'], round(s['GB/s'])) for s in d['stages']])
"
cat $OUT/kbench_real.log
