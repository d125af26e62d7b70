#!/bin/bash
# usage: tools/gpu_1024_variants.sh N -- 1024^3 c2c on N GPUs with the tile variants of the long strided / contiguous kernels
N=${1:-2}
OUT=gpurun_out/variants_1024_${N}gpu
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
show(){ grep -h '"metric"' $1 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'], '|', round(d['value'],1), 'GFlop/s', d['ms_per_step'], 'ms/step err', d['max_roundtrip_error'])
    for s in d['stages'][:len(d['stages'])//2]: print('   %-9s %-42s %8.4f ms sent %7.1f MB nvl %6.1f GB/s' % (s['direction'], s['stage'], s['ms'], s['sent_bytes']/1e6, s.get('nvlink_GB/s', 0)))
"; }
for prec in float double; do
for v in default 0; do
  if [ $v = default ]; then unset HEFFTE_B200_STRIDED_BIG; else export HEFFTE_B200_STRIDED_BIG=$v; fi
  timeout 200 $TR --master-port 29712 bench.py --gpus $N --size 1024 1024 1024 --precision $prec --steps 5 --warmup 3 \
      --no-secondary --no-e2e --no-cpu-baseline --no-parity > $OUT/bench_${prec}_$v.log 2>&1; echo "$prec variant $v rc=$?"
  show $OUT/bench_${prec}_$v.log
done
done
unset HEFFTE_B200_STRIDED_BIG
