#!/bin/bash
# Multi-GPU session on one N-GPU box: scaling benches, 2-D weak scaling, mrx rebalance, C++ NCCL host.
# Usage (under gpurun --gpus N):  bash tools/gpu_scale.sh <tag> <N>
set -u
TAG=${1:-s02}
N=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > "$OUT/gpu.txt" 2>&1
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29600
run() { # n outfile args...
  local n=$1 f=$2; shift 2
  port=$((port + 1))
  if [ "$n" = 1 ]; then timeout 900 python bench.py --gpus 1 "$@" > "$OUT/$f.json" 2> "$OUT/$f.err"
  else timeout 900 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n "$@" > "$OUT/$f.json" 2> "$OUT/$f.err"; fi
  echo "exit $?" >> "$OUT/$f.err"; tail -c 600 "$OUT/$f.json"; echo
}
run $N bench_n$N --steps 20 --warmup 5
for n in 4 2; do [ $n -lt $N ] && run $n bench_n$n --steps 20 --warmup 5 --no-e2e; done
run $N bench_cherenkov_n$N --workload cherenkov --steps 20 --warmup 5 --no-e2e
port=$((port + 1))
timeout 600 $TR --nproc-per-node $N --master-port $port tools/mrx_rebalance.py > "$OUT/mrx_rebalance_n$N.json" 2> "$OUT/mrx_rebalance.err"
echo "exit $?" >> "$OUT/mrx_rebalance.err"; tail -c 1500 "$OUT/mrx_rebalance_n$N.json"; echo
port=$((port + 1))
timeout 600 $TR --nproc-per-node $N --master-port $port --no-python host/host_nccl_demo 128 20 > "$OUT/host_nccl_n$N.txt" 2> "$OUT/host_nccl.err"
echo "exit $?" >> "$OUT/host_nccl.err"; tail -5 "$OUT/host_nccl_n$N.txt"
