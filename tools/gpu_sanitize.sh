#!/bin/bash
# compute-sanitizer (memcheck, racecheck, initcheck) on tools/sanitize_small.py
OUT=gpurun_out/${1:-san}; mkdir -p $OUT
F=$OUT/compute_sanitizer.txt
echo "# compute-sanitizer on tools/sanitize_small.py (3-D / 2-D / 1-D tiled kernels, lazy and eager sort, migration, halos, moments, host-buffer step, chunk move, far movers, boundary kinds, growing segments, round-1 kernel), B200" > $F
for tool in memcheck racecheck initcheck; do
  echo "== $tool" >> $F
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_small.py 2>&1 | grep -v "^=========     \|^========= *$" | grep -i "ok\|SUMMARY\|error\|hazard\|Invalid\|Uninit\|Traceback" | head -60 >> $F
done
cat $F
