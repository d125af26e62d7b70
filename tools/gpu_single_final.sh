#!/bin/bash
# 1 GPU, final tree: the GPU suite the way the driver runs it, smoke(), the default bench line, the reference arm
OUT=gpurun_out/r02_final
mkdir -p $OUT
(time python -m pytest tests -x -q -m gpu) > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -2 $OUT/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; tail -c 600 $OUT/bench_reference.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_final/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity_ok','gpu_launches')}, 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline']['value'], d['clocks'])
for s in d.get('secondary',[]): print('  ', {k:s.get(k) for k in ('workload','value','ms_per_step','parity_ok')})
PY
