#!/bin/bash
# Round-2 GPU session: parity tests, short benches of both row kernels, ncu of the new one.
# Usage (under gpurun, from the repo root):  bash tools/gpu_r02.sh <tag> [what...]
#   what: tests tests1 bench bench1 benchfull benchref benchwl launches full sweep   (default: tests bench bench1 full)
set -u
TAG=${1:-r02a}
shift || true
WHAT=${*:-tests bench bench1 full}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
has() { [[ " $WHAT " == *" $1 "* ]]; }

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
nproc > "$OUT/host_cores.txt"; grep -m1 "model name" /proc/cpuinfo >> "$OUT/host_cores.txt"

if has tests; then
  timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 ${PYTEST_ARGS:-} > "$OUT/pytest_gpu.log" 2>&1
  echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
  tail -25 "$OUT/pytest_gpu.log"
fi
if has tests1; then
  PICNIX_ROW_KERNEL=1 timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 > "$OUT/pytest_gpu_v1.log" 2>&1
  echo "pytest exit $?" >> "$OUT/pytest_gpu_v1.log"
  tail -5 "$OUT/pytest_gpu_v1.log"
fi
if has bench; then
  timeout 900 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu > "$OUT/bench_short.json" 2> "$OUT/bench_short.err"
  echo "bench exit $?" >> "$OUT/bench_short.err"
  cat "$OUT/bench_short.json"; tail -3 "$OUT/bench_short.err"
fi
if has bench1; then
  PICNIX_ROW_KERNEL=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-parity > "$OUT/bench_short_v1.json" 2> "$OUT/bench_short_v1.err"
  cat "$OUT/bench_short_v1.json"
fi
if has benchfull; then
  timeout 1500 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"
  echo "bench exit $?" >> "$OUT/bench.err"
  cat "$OUT/bench.json"; tail -3 "$OUT/bench.err"
fi
if has benchref; then
  timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"
  cat "$OUT/bench_ref.json"
fi
if has benchwl; then
  for w in ${WORKLOADS:-cherenkov twostream}; do
    timeout 900 python bench.py --workload $w --steps 20 --warmup 5 ${BENCHWL_ARGS:---e2e-steps 2} > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"
    echo "bench exit $?" >> "$OUT/bench_$w.err"
    cat "$OUT/bench_$w.json"; tail -2 "$OUT/bench_$w.err"
  done
fi
if has launches; then
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s ${LAUNCH_SKIP:-1030} -c 400 --csv \
    --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > "$OUT/launches_run.log" 2>&1
  python tools/summarize_launches.py "$OUT/launches.csv" > "$OUT/launches_summary.txt" 2>&1
  cat "$OUT/launches_summary.txt"
fi
if has full; then
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNEL:-row_push_kernel}" \
    -s ${NCU_SKIP:-3} -c 1 -f -o "$OUT/prof_${NCU_NAME:-row_push}" \
    python bench.py ${NCU_BENCH_ARGS:-} --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > "$OUT/full_run.log" 2>&1
  ls -la "$OUT"
fi
if has sweep; then
  timeout 1500 python tools/sweep.py > "$OUT/sweep.txt" 2>&1
  cat "$OUT/sweep.txt"
fi
