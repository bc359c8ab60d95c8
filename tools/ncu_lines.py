#!/usr/bin/env python
"""Join an ncu report's per-instruction samples with nvdisasm line info of the matching cubin.

    python tools/ncu_lines.py <report.ncu-rep> <object.o|.so> <kernel-name-substring> [top N]

Prints the headline metrics, the opcode mix, stall reasons and the hottest source lines.
The object must be the build that was profiled (same SASS).
"""
import csv
import io
import os
import re
import subprocess
import sys
from collections import Counter, defaultdict


def run(cmd):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep, obj, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    row = [r for r in raw[2:] if kname in r[hdr.index("Kernel Name")]][0]
    want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
            "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
            "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed_pipe_lsu.sum", "sm__cycles_elapsed.avg"]
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w} [{units[i]}] = {row[i]}")

    src = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name",
                                           "regex:" + kname]))))
    h = next(r for r in src if "Source" in r and "# Samples" in r)
    data = [r for r in src[src.index(h) + 1:] if len(r) == len(h)]
    iS, iI, isrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    tot, ti = sum(int(r[iS]) for r in data), sum(int(r[iI]) for r in data)
    print(f"\nSASS instructions {len(data)}, samples {tot}, warp instructions executed {ti}")
    ops, ops_s = Counter(), Counter()
    for r in data:
        f = r[isrc].split()
        op = (f[1] if f[0].startswith("@") else f[0]).split(".")[0]
        ops[op] += int(r[iI])
        ops_s[op] += int(r[iS])
    print("opcode mix (executed % / samples %): " + ", ".join(
        f"{op} {100 * n / ti:.1f}/{100 * ops_s[op] / tot:.1f}" for op, n in ops.most_common(16)))
    stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    st = {c: sum(int(r[h.index(c)]) for r in data) for c in stalls}
    print("stalls %: " + ", ".join(f"{c[6:]} {100 * v / tot:.1f}" for c, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]))

    # line info from the object
    tmp = "/tmp/ncu_lines_cubin"
    os.makedirs(tmp, exist_ok=True)
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    locs = None
    for f in os.listdir(tmp):
        text = run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)])
        sections = re.split(r"\n//-+ \.text\.", text)
        for sec in sections:
            if kname in sec.split("\n", 1)[0] and "row_kernelILb1ELi0ELi0" in sec.split("\n", 1)[0] or \
                    (kname in sec.split("\n", 1)[0] and "row_kernel" not in kname):
                cur, out = None, []
                for ln in sec.split("\n"):
                    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
                    if m:
                        cur = (os.path.basename(m.group(1)), int(m.group(2)))
                    elif re.match(r"\s*/\*[0-9a-f]+\*/", ln):
                        out.append(cur)
                if len(out) == len(data):
                    locs = out
    if locs is None:
        print("no matching SASS section found in", obj)
        return
    agg = defaultdict(lambda: [0, 0, 0, 0])
    iW = h.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in h else None
    iWi = h.index("L1 Wavefronts Shared Ideal") if "L1 Wavefronts Shared Ideal" in h else None
    for loc, r in zip(locs, data):
        k = loc or ("?", 0)
        agg[k][0] += int(r[iS])
        agg[k][1] += int(r[iI])
        if iW is not None:
            agg[k][2] += int(r[iW] or 0)
            agg[k][3] += int(r[iWi] or 0)
    tw = sum(v[2] for v in agg.values())
    if tw:
        print(f"\nshared-memory wavefronts {tw} (ideal {sum(v[3] for v in agg.values())}); by source line (% of all, "
              f"wavefronts per warp instruction):")
        for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:14]:
            print(f"  {f}:{l:4d} {100 * v[2] / tw:5.1f}%  ideal {100 * v[3] / tw:5.1f}%")
    srcdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "picnix_b200", "csrc")
    cache = {}
    print("\nhottest source lines (samples % / executed %):")
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in cache:
            p = os.path.join(srcdir, f)
            cache[f] = open(p).read().split("\n") if os.path.exists(p) else []
        text = cache[f][l - 1].strip()[:88] if 0 < l <= len(cache[f]) else ""
        print(f"{f}:{l:4d} {100 * v[0] / tot:5.2f} / {100 * v[1] / ti:5.2f}  {text}")


if __name__ == "__main__":
    main()
