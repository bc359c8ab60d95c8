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

# round 2: the 2-D and 1-D tiled kernels, physical boundary kinds, growing segments, the round-1 kernel
from picnix_b200 import capi  # noqa: E402

# 2-D Harris sheet between conducting walls (rowpush2d.cu, boundary.cu)
nd, cd = (1, 32, 32), (1, 2, 2)
sim = CudaSim(nd, cd, Ns=2, cc=1.0, delh=0.2, order=2, periodic=(1, 0, 1))
for side in (0, 1):
    sim.set_boundary_condition(1, side, capi.BC_CONDUCTING)
problems.setup_harris_sheet(sim, nd, cd, delh=0.2, seed=1, ncs=8, nbg=4)
sim.step(0.1, 4)
sim.deposit_moment()
sim.exchange(2)
sim.synchronize()
print("mrx 2-D ok", int(sim.get_np_all().sum()))
sim.close()

# 1-D two-stream in 8-cell chunks (rowpush1d.cu) and the shock tube's wall + inflow
nd, cd = (1, 1, 64), (1, 1, 8)
sim = CudaSim(nd, cd, Ns=3, cc=50.0, delh=1.0, order=2)
problems.setup_uniform_plasma(sim, nd, cd, problems.TWOSTREAM_SPECIES, (8, 8, 16), B0=(10.0, 0, 0), seed=2)
sim.step(0.01, 4)
sim.step(0.15, 1)  # far movers
sim.synchronize()
print("two-stream 1-D ok", int(sim.get_np_all().sum()))
sim.close()
nd, cd = (1, 1, 64), (1, 1, 4)
sim = CudaSim(nd, cd, Ns=2, cc=1.0, delh=1.0, order=2, periodic=(1, 1, 0))
sim.set_boundary_condition(2, 0, capi.BC_WALL)
sim.set_boundary_condition(2, 1, capi.BC_INFLOW, [0, 0, 0, 0.0, 0.1, 0.0])
species = [dict(qm=-1.0, ro=1.0, vt=0.1, drift=(-0.1, 0.0, 0.0)), dict(qm=+0.04, ro=25.0, vt=0.02, drift=(-0.1, 0.0, 0.0))]
problems.setup_uniform_plasma(sim, nd, cd, species, (8, 8), B0=(0.0, 0.1, 0.0), seed=3)
sim.step(0.5, 6)
sim.synchronize()
print("shock 1-D ok", int(sim.get_np_all().sum()))
sim.close()

# growth: everything converges on one chunk (spill list, re-layout), both check modes
nd, cd = (16, 16, 16), (2, 2, 2)
for always in (0, 1):
    sim = CudaSim(nd, cd, Ns=2, cc=1.0, delh=1.0, order=2)
    sim.set_option("check_growth", always)
    dims = problems.chunk_dims(nd, cd)
    _, coord = sim.chunkmap()
    for isp, (q, m) in enumerate(problems.species_charge_mass([dict(qm=-1.0, ro=1e-3), dict(qm=0.1, ro=1e-2)], (4, 4))):
        sim.set_species(isp, q, m)
    for ic in range(sim.nchunk):
        sim.set_field(ic, capi.FIELD_UF, np.zeros(sim.shape + (6,)))
        parts = problems.make_chunk_particles(ic, coord[ic], dims, 1.0, [dict(vt=0.0), dict(vt=0.0)], (4, 4), 11)
        for isp, xu in enumerate(parts):
            xu[:, 3:6] = -0.1 * (xu[:, 0:3] - 4.0)
            sim.set_particles(ic, isp, xu)
    sim.finalize_setup()
    sim.step(0.4, 30)
    sim.synchronize()
    print("growth ok", always, int(sim.get_np_all().sum()), int(sim.get_np_all().max()), sim.growth_stats())
    sim.close()

# the round-1 kernel kept under experiments/
sim = CudaSim((16, 16, 16), (2, 2, 2), Ns=2, cc=10.0, delh=1.0, order=2)
sim.set_option("row_kernel", 1)
problems.setup_uniform_plasma(sim, (16, 16, 16), (2, 2, 2), problems.THERMAL_SPECIES, (4, 4), B0=(5.0, 0, 0), seed=2)
sim.step(0.05, 2)
sim.synchronize()
print("round-1 kernel ok", int(sim.get_np_all().sum()))
sim.close()
