#!/bin/bash
# usage: tools/gpu_multi_probe.sh N   -- multi-rank parity worker + bench variants on N GPUs of one box (run under gpurun --gpus N)
N=${1:-8}
OUT=gpurun_out/multi_${N}gpu
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 420 $TR --master-port 29701 tests/multi_rank_worker.py --quick > $OUT/parity.log 2>&1; echo "parity rc=$?" | tee -a $OUT/parity.log
timeout 240 $TR --master-port 29702 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_c2c_f64_512.log 2>&1; echo "bench rc=$?"
timeout 200 $TR --master-port 29703 bench.py --gpus $N --steps 10 --warmup 3 --io-pencils --no-e2e > $OUT/bench_c2c_f64_512_iopencils.log 2>&1; echo "bench-io rc=$?"
HEFFTE_B200_DISABLE_P2P=1 timeout 200 $TR --master-port 29704 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > $OUT/bench_c2c_f64_512_nccl.log 2>&1; echo "bench-nccl rc=$?"
timeout 200 $TR --master-port 29705 bench.py --gpus $N --steps 10 --warmup 3 --kind r2c --no-e2e > $OUT/bench_r2c_f64_512.log 2>&1; echo "bench-r2c rc=$?"
timeout 240 $TR --master-port 29706 bench.py --gpus $N --steps 5 --warmup 3 --size 1024 1024 1024 --precision float --reorder --no-e2e > $OUT/bench_c2c_f32_1024_reorder_pencils.log 2>&1; echo "bench-1024 rc=$?"
timeout 240 $TR --master-port 29707 bench.py --gpus $N --steps 5 --warmup 3 --size 1024 1024 1024 --precision float --reorder --slabs --no-e2e > $OUT/bench_c2c_f32_1024_reorder_slabs.log 2>&1; echo "bench-1024-slabs rc=$?"
grep -h '"metric"' $OUT/bench_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'], '|', d['config']['comm'], '|', round(d['value'],1), 'GFlop/s', d['ms_per_step'], 'ms/step err', d['max_roundtrip_error'])
"
tail -3 $OUT/parity.log
