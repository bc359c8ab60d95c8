#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck):
3-D thermal plasma through the tiled fused kernel (lazy and eager sort), migration, halos, moments,
the host-buffer step and a chunk move.

    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from picnix_b200 import CudaSim, problems  # noqa: E402

nd, cd = (16, 16, 16), (2, 2, 2)
for lazy in (1, 0):
    sim = CudaSim(nd, cd, Ns=2, cc=10.0, delh=1.0, order=2)
    sim.set_option("lazy_sort", lazy)
    problems.setup_uniform_plasma(sim, nd, cd, problems.THERMAL_SPECIES, (4, 4), B0=(5.0, 0, 0), seed=2, perturb=0.01)
    sim.step(0.05, 3)
    sim.deposit_moment()
    sim.exchange(2)
    e = sim.get_energy().sum()
    st = sim.host_state(pinned=False)
    sim.step_host(st, 0.05, 1)
    buf = sim.chunk_pack(1)
    sim.chunk_unpack(1, buf)
    sim.sort_particle()
    sim.step(0.2, 1)  # large step: far movers, many migrants
    sim.synchronize()
    print("lazy" if lazy else "eager", "ok", int(sim.get_np_all().sum()), float(e))
    sim.close()
