#!/bin/bash
# usage: tools/gpu_1024_variants.sh N -- 1024^3 fp32 c2c on N GPUs with the tile variants of the 1024-point strided kernel
N=${1:-2}
OUT=gpurun_out/variants_1024_${N}gpu
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for v in 0 1 2; do
  HEFFTE_B200_STRIDED_1024=$v timeout 200 $TR --master-port 2971$v bench.py --gpus $N --size 1024 1024 1024 --precision float --steps 5 --warmup 3 \
      --no-secondary --no-e2e --no-cpu-baseline --no-parity > $OUT/bench_v$v.log 2>&1; echo "variant $v rc=$?"
  grep -h '"metric"' $OUT/bench_v$v.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'], '|', round(d['value'],1), 'GFlop/s', d['ms_per_step'], 'ms/step err', d['max_roundtrip_error'])
    for s in d['stages'][:len(d['stages'])//2]: print('   %-9s %-42s %8.4f ms sent %7.1f MB nvl %6.1f GB/s' % (s['direction'], s['stage'], s['ms'], s['sent_bytes']/1e6, s.get('nvlink_GB/s', 0)))
"
done
