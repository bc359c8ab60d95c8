#!/usr/bin/env python
"""Time every phase of the step separately on one GPU (CUDA events on the arena's stream)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from picnix_b200 import problems  # noqa: E402
from picnix_b200.distributed import DistributedSim  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=128)
ap.add_argument("--ppc", type=int, default=32)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--order", type=int, default=2)
args = ap.parse_args()

nd = (args.cells,) * 3
cd = tuple(n // 16 for n in nd)
sim = DistributedSim(nd, cd, Ns=2, cc=10.0, order=args.order)
stream = torch.cuda.Stream()
sim.set_stream(stream.cuda_stream)
problems.setup_uniform_plasma(sim, nd, cd, problems.THERMAL_SPECIES, (args.ppc, args.ppc), B0=(5.0, 0, 0), seed=1)
npart = int(sim.get_np_all().sum())
dt = 0.05
phases = [
    ("push_bfd", lambda: sim.push_bfd(0.5 * dt)),
    ("push_velocity", lambda: sim.push_velocity(dt)),
    ("push_position", lambda: sim.push_position(dt)),
    ("deposit_current", lambda: sim.deposit_current(dt)),
    ("halo_cur", lambda: sim.exchange(1)),
    ("migrate", lambda: sim.boundary_begin(3)),
    ("push_bfd2", lambda: sim.push_bfd(0.5 * dt)),
    ("push_efd", lambda: sim.push_efd(dt)),
    ("halo_emf", lambda: sim.exchange(0)),
    ("sort", lambda: sim.boundary_end(3)),
    # diagnostics cadence (not part of the step total below)
    ("deposit_moment*", lambda: sim.deposit_moment()),
    ("halo_mom*", lambda: sim.exchange(2)),
]
acc = {name: [] for name, _ in phases}
with torch.cuda.stream(stream):
    for rep in range(args.reps + 2):
        for name, fn in phases:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            if rep >= 2:
                acc[name].append(e0.elapsed_time(e1))
sim.synchronize()
total = 0.0
for name, _ in phases:
    ms = float(np.mean(acc[name]))
    if not name.endswith("*"):
        total += ms
    print(f"{name:18s} {ms:9.3f} ms   {npart / ms / 1e6:10.1f} Mparticles/ms-equivalent" if False else
          f"{name:18s} {ms:9.3f} ms   {ms * 1e6 / npart:8.3f} ns/particle")
print(f"{'total':18s} {total:9.3f} ms   -> {npart / total * 1e3:.3e} particle-steps/s  (np={npart})")
