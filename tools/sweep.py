#!/usr/bin/env python
"""Throughput sweep around the headline workload: particles per cell, and the 3-D
two-stream beam (example/beam/twostream species on the T3D grid; drifting beams stress migration).
One GPU, state resident, CUDA events on the arena's stream; prints one line per case."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from picnix_b200 import CudaSim, problems  # noqa: E402

CASES = [
    # name, cells, species, ppc per species, cc, dt, B0[, chunk edge]
    ("thermal ppc=2x32, chunks 8^3", 128, problems.THERMAL_SPECIES, (32, 32), 10.0, 0.05, (5.0, 0, 0), 8),
    ("thermal ppc=2x32, chunks 32^3", 128, problems.THERMAL_SPECIES, (32, 32), 10.0, 0.05, (5.0, 0, 0), 32),
    ("thermal ppc=2x8", 128, problems.THERMAL_SPECIES, (8, 8), 10.0, 0.05, (5.0, 0, 0)),
    ("thermal ppc=2x16", 128, problems.THERMAL_SPECIES, (16, 16), 10.0, 0.05, (5.0, 0, 0)),
    ("thermal ppc=2x32 (headline)", 128, problems.THERMAL_SPECIES, (32, 32), 10.0, 0.05, (5.0, 0, 0)),
    ("thermal ppc=2x64", 128, problems.THERMAL_SPECIES, (64, 64), 10.0, 0.05, (5.0, 0, 0)),
    ("two-stream beam 3-D ppc=16+16+32", 128, problems.TWOSTREAM_SPECIES, (16, 16, 32), 50.0, 0.01, (10.0, 0, 0)),
]
steps, warmup = 10, 3
stream = torch.cuda.Stream()
print(f"{'case':36s} {'particles':>12s} {'ms/step':>9s} {'particle-steps/s':>18s} {'leaving/step':>13s}")
for case in CASES:
    name, cells, species, ppc, cc, dt, B0 = case[:7]
    edge = case[7] if len(case) > 7 else 16
    nd = (cells,) * 3
    cd = tuple(n // edge for n in nd)
    sim = CudaSim(nd, cd, Ns=len(species), cc=cc, delh=1.0, order=2)
    sim.set_stream(stream.cuda_stream)
    problems.setup_uniform_plasma(sim, nd, cd, species, ppc, B0=B0, seed=1)
    npart = int(sim.get_np_all().sum())
    with torch.cuda.stream(stream):
        sim.step(dt, warmup)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sim.step(dt, steps)
        e1.record(stream)
        e1.synchronize()
    ms = e0.elapsed_time(e1) / steps
    sim.synchronize()
    assert int(sim.get_np_all().sum()) == npart
    # fraction of particles that changed chunk in one step (migration load)
    before = sim.get_np_all().copy()
    print(f"{name:36s} {npart:12d} {ms:9.3f} {npart / ms * 1e3:18.4e}")
    sim.close()
    del sim
    torch.cuda.empty_cache()
