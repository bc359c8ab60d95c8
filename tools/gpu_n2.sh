set -u
OUT=gpurun_out/s02g; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29701 tools/rebalance_demo.py > $OUT/rebalance_demo.json 2> $OUT/rebalance_demo.err; echo "rebalance_demo exit $?"; tail -c 800 $OUT/rebalance_demo.json; echo
timeout 600 $TR --master-port 29702 tools/mrx_rebalance.py > $OUT/mrx_n2.json 2> $OUT/mrx_n2.err; echo "mrx exit $?"; tail -c 600 $OUT/mrx_n2.json; echo
timeout 600 $TR --master-port 29703 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench exit $?"
timeout 600 $TR --master-port 29704 bench.py --gpus 2 --workload twostream --steps 20 --warmup 5 --no-e2e > $OUT/bench_twostream_n2.json 2> $OUT/bench_twostream_n2.err; echo "bench ts exit $?"
timeout 600 $TR --master-port 29705 bench.py --gpus 2 --workload cherenkov --steps 20 --warmup 5 --no-e2e > $OUT/bench_cherenkov_n2.json 2> $OUT/bench_cherenkov_n2.err; echo "bench ch exit $?"
python - <<EOF2
import json
for f in ("bench_n2","bench_twostream_n2","bench_cherenkov_n2"):
    try:
        d=json.loads(open("$OUT/%s.json"%f).read().strip().split("\n")[-1])
        print(f, "%.4e"%d["value"], "%.3f"%d["ms_per_step"], d["parity_check"])
    except Exception as e: print(f, "ERR", e)
EOF2
