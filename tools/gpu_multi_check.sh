#!/bin/bash
# usage: tools/gpu_multi_check.sh N -- short check on N GPUs: multi-rank parity (quick) + the headline bench line
N=${1:-4}
OUT=gpurun_out/multi_${N}gpu_check
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29701 tests/multi_rank_worker.py --quick --subcomm > $OUT/parity.log 2>&1; echo "parity rc=$?" | tee -a $OUT/parity.log
timeout 240 $TR --master-port 29702 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_c2c_f64_512.log 2>&1; echo "bench rc=$?"
grep -h '"metric"' $OUT/bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'], '|', d['config']['comm'], '|', round(d['value'],1), 'GFlop/s', d['ms_per_step'], 'ms/step err', d['max_roundtrip_error'])
    for s in d['stages']: print('   %-9s %-42s %8.4f ms sent %7.1f MB nvl %6.1f GB/s' % (s['direction'], s['stage'], s['ms'], s['sent_bytes']/1e6, s.get('nvlink_GB/s', 0)))
"
tail -3 $OUT/parity.log
