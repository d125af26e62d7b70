#!/bin/bash
# usage: tools/gpu_multi_probe.sh N   -- multi-rank parity worker + bench variants on N GPUs of one box (run under gpurun --gpus N)
N=${1:-8}
OUT=gpurun_out/multi_${N}gpu_b
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
B="--gpus $N --warmup 3 --no-e2e"
timeout 420 $TR --master-port 29701 tests/multi_rank_worker.py --quick > $OUT/parity.log 2>&1; echo "parity rc=$?" | tee -a $OUT/parity.log
timeout 240 $TR --master-port 29702 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_c2c_f64_512.log 2>&1; echo "bench rc=$?"
HEFFTE_B200_DECOMPOSITION=pencils timeout 200 $TR --master-port 29703 bench.py $B --steps 10 > $OUT/bench_c2c_f64_512_pencils.log 2>&1; echo "bench-pencils rc=$?"
timeout 200 $TR --master-port 29705 bench.py $B --steps 10 --kind r2c > $OUT/bench_r2c_f64_512.log 2>&1; echo "bench-r2c rc=$?"
timeout 200 $TR --master-port 29706 bench.py $B --steps 10 --kind r2r > $OUT/bench_r2r_f64_512.log 2>&1; echo "bench-r2r rc=$?"
timeout 200 $TR --master-port 29707 bench.py $B --steps 10 --kind conv > $OUT/bench_conv_f64_512.log 2>&1; echo "bench-conv rc=$?"
timeout 240 $TR --master-port 29708 bench.py $B --steps 5 --size 1024 1024 1024 --precision float --reorder > $OUT/bench_c2c_f32_1024.log 2>&1; echo "bench-1024 rc=$?"
grep -h '"metric"' $OUT/bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'], '|', d['config']['comm'], '|', round(d['value'],1), 'GFlop/s', d['ms_per_step'], 'ms/step err', d['max_roundtrip_error'])
"
tail -3 $OUT/parity.log
