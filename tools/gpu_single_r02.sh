#!/bin/bash
# 1 GPU: the whole GPU suite the way the driver runs it, the default bench line, launch list + full captures of the new kernels
OUT=gpurun_out/r02g
mkdir -p $OUT
(time python -m pytest tests -x -q -m gpu) > $OUT/pytest.log 2>&1; tail -5 $OUT/pytest.log
python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -2 $OUT/bench.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for v in 0 1; do HEFFTE_B200_STRIDED_BIG=$v python bench.py --size 1024 1024 1024 --precision float --steps 5 --warmup 3 --no-secondary --no-e2e --no-cpu-baseline --no-parity > $OUT/bench_1024_f32_big$v.json 2>/dev/null; python -c "
import json,sys
d=json.loads(open('$OUT/bench_1024_f32_big$v.json').read().strip().splitlines()[-1]); print('1024^3 fp32 STRIDED_BIG=$v', round(d['value'],1), [ (s['dim'], round(s['GB/s'])) for s in d['stages'] if s['direction']=='forward'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c2c.csv python bench.py --steps 2 --warmup 1 --no-secondary --no-parity --no-e2e --no-cpu-baseline > $OUT/ncu_c2c.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_conv.csv python bench.py --kind conv --steps 2 --warmup 1 --no-secondary --no-parity --no-e2e --no-cpu-baseline > $OUT/ncu_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_contig_real2 -c 2 -o $OUT/ncu_contig_real2 python bench.py --kind r2c --steps 1 --warmup 1 --no-secondary --no-parity --no-e2e --no-cpu-baseline > $OUT/ncu_full_r2c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_strided_real2 -c 2 -o $OUT/ncu_strided_real2 python bench.py --kind r2r --steps 1 --warmup 1 --no-secondary --no-parity --no-e2e --no-cpu-baseline > $OUT/ncu_full_r2r.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_strided_conv -c 1 -o $OUT/ncu_strided_conv python bench.py --kind conv --steps 1 --warmup 1 --no-secondary --no-parity --no-e2e --no-cpu-baseline > $OUT/ncu_full_conv.log 2>&1
ls -la $OUT
