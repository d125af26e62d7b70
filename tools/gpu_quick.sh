#!/bin/bash
# 1 GPU: 1024^3 with the tile variants of the long strided kernel on the LOCAL passes
OUT=gpurun_out/r02h
mkdir -p $OUT
for prec in float double; do for v in 0 1 2; do HEFFTE_B200_STRIDED_BIG=$v python bench.py --size 1024 1024 1024 --precision $prec --steps 5 --warmup 3 --no-secondary --no-e2e --no-cpu-baseline --no-parity > $OUT/bench_1024_${prec}_big$v.json 2>/dev/null; python -c "
import json,sys
d=json.loads(open('$OUT/bench_1024_${prec}_big$v.json').read().strip().splitlines()[-1]); print('1024^3 $prec STRIDED_BIG=$v', round(d['value'],1), [ (s['dim'], s['direction'][0], round(s['GB/s'])) for s in d['stages']])"; done; done
