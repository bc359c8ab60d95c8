#!/bin/bash
# 8-GPU re-measurement after the exchange merges: thermal3d N=1 and N=8 back to back on the same box,
# cherenkov and two-stream at N=8, C++ host at N=8, per-phase times at N=8.
set -u
TAG=${1:-s02n}; N=${2:-8}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python bench.py --steps 30 --warmup 8 --no-e2e --no-cpu > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"
timeout 900 $TR --nproc-per-node $N --master-port 29801 bench.py --gpus $N --steps 30 --warmup 8 > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"
timeout 900 $TR --nproc-per-node 4 --master-port 29806 bench.py --gpus 4 --steps 30 --warmup 8 --no-e2e > "$OUT/bench_n4.json" 2> "$OUT/bench_n4.err"
timeout 600 $TR --nproc-per-node $N --master-port 29802 bench.py --gpus $N --workload cherenkov --steps 30 --warmup 8 --no-e2e > "$OUT/bench_cherenkov_n$N.json" 2> "$OUT/bench_cherenkov_n$N.err"
timeout 600 $TR --nproc-per-node $N --master-port 29803 bench.py --gpus $N --workload twostream --steps 30 --warmup 8 --no-e2e > "$OUT/bench_twostream_n$N.json" 2> "$OUT/bench_twostream_n$N.err"
timeout 600 $TR --nproc-per-node $N --master-port 29804 --no-python host/host_nccl_demo 128 30 > "$OUT/host_nccl_n$N.txt" 2> "$OUT/host_nccl.err"
timeout 600 $TR --nproc-per-node $N --master-port 29805 tools/phase_times_multirank.py 2>&1 | grep -v "^\*\|OMP_NUM\|NCCL version" > "$OUT/phase_times_n$N.txt"
python - <<EOF2
import json,glob
for f in sorted(glob.glob("$OUT/bench*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f.split("/")[-1], "%.4e"%d["value"], "%.3f ms"%d["ms_per_step"], "parity", d["parity_check"] and d["parity_check"].get("ok"), "e2e", d["e2e"].get("value"))
    except Exception as e: print(f, "ERR", e)
EOF2
tail -2 "$OUT/host_nccl_n$N.txt"; cat "$OUT/phase_times_n$N.txt"
