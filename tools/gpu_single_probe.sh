#!/bin/bash
# 1-GPU round: GPU parity suite, smoke, bench lines of the BASELINE configurations, ncu launch list + full captures.
OUT=gpurun_out/single
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench_c2c_f64_512.log 2>&1; echo "bench rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --kind r2c --no-e2e --no-cpu-baseline > $OUT/bench_r2c_f64_512.log 2>&1; echo "bench r2c rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --kind r2r --no-e2e --no-cpu-baseline > $OUT/bench_r2r_f64_512.log 2>&1; echo "bench r2r rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --size 256 256 256 --precision float --no-e2e --no-cpu-baseline > $OUT/bench_c2c_f32_256.log 2>&1; echo "bench cfg2 rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --size 256 256 256 --precision float --kind r2r --no-e2e --no-cpu-baseline > $OUT/bench_r2r_f32_256.log 2>&1; echo "bench r2r f32 rc=$?"
# launch list of the headline bench (cold-cache serialised times: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench_1gpu.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
# full captures: the three kernels of one forward transform of each kind (first launches of the warm-up)
for kind in c2c r2c r2r; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_ -c 3 -o $OUT/ncu_${kind}_f64_512 -f \
      python bench.py --steps 1 --warmup 3 --kind $kind --no-e2e --no-cpu-baseline > $OUT/ncu_${kind}.log 2>&1; echo "ncu $kind rc=$?"
done
grep -h '"metric"' $OUT/bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'], '|', round(d['value'],1), 'GFlop/s', round(d['ms_per_step'],4), 'ms/step err', d['max_roundtrip_error'], 'roof', d['roofline'] and round(d['roofline']['frac'],3), [(s['kernel'], round(s['GB/s'])) for s in d['stages']])
"
ls -la $OUT
