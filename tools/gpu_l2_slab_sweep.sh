#!/bin/bash
# Next measurement to take (1 GPU): pairs of local transforms run slab by slab through the L2 cache (HEFFTE_B200_L2_SLAB_MB,
# csrc/transform.cpp l2_slab_planes): sweep of the slab size on the headline problem and on 256^3 fp32.
OUT=gpurun_out/l2_slab
mkdir -p $OUT
for mb in 0 8 16 32 48 64 96; do
  timeout 200 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --l2-slab-mb $mb > $OUT/c2c_f64_512_slab$mb.log 2>&1
  timeout 200 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --size 256 256 256 --precision float --l2-slab-mb $mb > $OUT/c2c_f32_256_slab$mb.log 2>&1
done
grep -h '"metric"' $OUT/*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:40], 'slab MB', d['config']['l2_slab_mb'], '|', round(d['value'],1), 'GFlop/s', round(d['ms_per_step'],4), 'ms/step')
"
