#!/bin/bash
# usage: tools/gpu_multi_registered.sh N [full] -- parity worker, then the headline with and without registered arrays; "full": the default line too
N=${1:-2}
OUT=gpurun_out/multi_${N}gpu_registered
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
show(){ grep -h '"metric"' $1 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'], '| registered', d['config'].get('registered_arrays'), '|', round(d['value'],1), 'GFlop/s', d['ms_per_step'], 'ms/step err', d['max_roundtrip_error'], 'parity', d.get('parity_rel_l2'))
    for s in d['stages'][:len(d['stages'])//2]: print('   %-9s %-42s %8.4f ms sent %7.1f MB nvl %6.1f GB/s' % (s['direction'], s['stage'], s['ms'], s['sent_bytes']/1e6, s.get('nvlink_GB/s', 0)))
    for x in d.get('secondary', []): print('   secondary:', x['workload'], x.get('layout','')[:40], round(x['value'],1), x['ms_per_step'], 'parity', x.get('parity_rel_l2'))
"; }
timeout 420 $TR --master-port 29701 tests/multi_rank_worker.py --quick --subcomm > $OUT/parity.log 2>&1; echo "parity rc=$?" | tee -a $OUT/parity.log; tail -2 $OUT/parity.log
timeout 200 $TR --master-port 29721 bench.py --gpus $N --steps 10 --warmup 3 --no-secondary --no-e2e --no-cpu-baseline > $OUT/bench_registered.log 2>&1; echo "registered rc=$?"; show $OUT/bench_registered.log
timeout 200 $TR --master-port 29722 bench.py --gpus $N --steps 10 --warmup 3 --no-secondary --no-e2e --no-cpu-baseline --no-register > $OUT/bench_not_registered.log 2>&1; echo "not registered rc=$?"; show $OUT/bench_not_registered.log
if [ "$2" = full ]; then timeout 400 $TR --master-port 29702 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_default.log 2>&1; echo "bench rc=$?"; show $OUT/bench_default.log; fi
