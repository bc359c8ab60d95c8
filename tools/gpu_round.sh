#!/bin/bash
# One GPU session: parity tests, smoke, bench, per-phase timing, ncu launch list, ncu full capture.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag] [what...]
#   what: tests smoke bench phases launches full   (default: all)
set -u
TAG=${1:-r01}
shift || true
WHAT=${*:-tests smoke bench phases launches full}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
has() { [[ " $WHAT " == *" $1 "* ]]; }

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
nproc > "$OUT/host_cores.txt"; grep -m1 "model name" /proc/cpuinfo >> "$OUT/host_cores.txt"

if has tests; then
  timeout 1500 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1
  echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
  tail -5 "$OUT/pytest_gpu.log"
fi
if has smoke; then
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
  echo "smoke exit $?" >> "$OUT/smoke.log"
  tail -3 "$OUT/smoke.log"
fi
if has bench; then
  timeout 1500 python bench.py --steps 20 --warmup 5 > "$OUT/bench.json" 2> "$OUT/bench.err"
  echo "bench exit $?" >> "$OUT/bench.err"
  cat "$OUT/bench.json"; tail -3 "$OUT/bench.err"
fi
if has benchref; then
  timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"
  cat "$OUT/bench_ref.json"
fi
if has phases; then
  timeout 900 python tools/phase_times.py > "$OUT/phases.txt" 2>&1
  cat "$OUT/phases.txt"
fi
if has launches; then
  # the set-up uploads launch 1024 transposes + 6 set-up kernels; skip them and list 4 whole steps
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s ${LAUNCH_SKIP:-1030} -c 400 --csv \
    --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > "$OUT/launches_run.log" 2>&1
  python tools/summarize_launches.py "$OUT/launches.csv" > "$OUT/launches_summary.txt" 2>&1
  cat "$OUT/launches_summary.txt"
fi
if has full; then
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNEL:-row_kernel}" \
    -s ${NCU_SKIP:-3} -c 1 -f -o "$OUT/prof_${NCU_NAME:-row_kernel}" \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > "$OUT/full_run.log" 2>&1
  ls -la "$OUT"
fi
