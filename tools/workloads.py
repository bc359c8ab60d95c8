#!/usr/bin/env python
"""Throughput of the BASELINE.json configurations in their NATIVE dimension on one GPU (state resident,
CUDA events on the arena's stream): the reference's example problems scaled up in extent until they fill
a B200, everything else (chunk shape, species, particles per cell, time step, boundary conditions) as
shipped in example/*/config.toml.

    python tools/workloads.py [name ...]        names: thermal3d twostream cherenkov mrx shock

One JSON line per workload: ms per step, particle-steps/s, the share of the fused push+deposit kernel,
and its algorithmic HBM fraction (120 B per particle + 112 B per cell).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from picnix_b200 import CudaSim, capi, problems  # noqa: E402

CHERENKOV_SPECIES = [dict(qm=-1.0, ro=1.0, vt=0.1, drift=(0.1, 0.0, 0.0)),
                     dict(qm=+1.0, ro=1.0, vt=0.1, drift=(0.1, 0.0, 0.0))]
SHOCK_SPECIES = [dict(qm=-1.0, ro=1.0, vt=0.1, drift=(-0.1, 0.0, 0.0)),
                 dict(qm=+0.04, ro=25.0, vt=0.02, drift=(-0.1, 0.0, 0.0))]


def build(name):
    if name == "thermal3d":      # example/thermal with Nz, Ny > 1 (the headline)
        nd, cd = (128,) * 3, (8,) * 3
        sim = CudaSim(nd, cd, Ns=2, cc=10.0, delh=1.0, order=2)
        problems.setup_uniform_plasma(sim, nd, cd, problems.THERMAL_SPECIES, (32, 32), B0=(5.0, 0, 0), seed=1)
        return sim, 0.05, "thermal 3-D 128^3, 16^3 chunks, 2x32 ppc"
    if name == "twostream":      # example/beam/twostream/config.toml: 8-cell chunks, 16+16+32 ppc
        nd, cd = (1, 1, 1 << 18), (1, 1, 1 << 15)
        sim = CudaSim(nd, cd, Ns=3, cc=50.0, delh=1.0, order=2)
        problems.setup_uniform_plasma(sim, nd, cd, problems.TWOSTREAM_SPECIES, (16, 16, 32), B0=(10.0, 0, 0), seed=1)
        return sim, 0.01, "two-stream 1-D Nx=2^18, 8-cell chunks, 16+16+32 ppc"
    if name == "cherenkov":      # example/cherenkov/config.toml: 16^2 chunks, 2x32 ppc, drifting pair plasma
        nd, cd = (1, 1024, 1024), (1, 64, 64)
        sim = CudaSim(nd, cd, Ns=2, cc=1.0, delh=0.1, order=2)
        problems.setup_uniform_plasma(sim, nd, cd, CHERENKOV_SPECIES, (32, 32), delh=0.1, seed=1)
        return sim, 0.05, "cherenkov 2-D 1024^2, 16^2 chunks, 2x32 ppc, u0 = 0.1"
    if name == "mrx":            # example/mrx/config.toml: Harris sheet between conducting walls
        nd, cd = (1, 1024, 1024), (1, 64, 64)
        sim = CudaSim(nd, cd, Ns=2, cc=1.0, delh=0.2, order=2, periodic=(1, 0, 1))
        for side in (0, 1):
            sim.set_boundary_condition(1, side, capi.BC_CONDUCTING)
        problems.setup_harris_sheet(sim, nd, cd, delh=0.2, seed=1)
        return sim, 0.1, "mrx 2-D Harris sheet 1024^2, 16^2 chunks, ncs=50 nbg=10, conducting walls"
    if name == "shock":          # example/shock/config.toml: 16-cell chunks, 2x32 ppc, wall + inflow fields
        nd, cd = (1, 1, 1 << 18), (1, 1, 1 << 14)
        sim = CudaSim(nd, cd, Ns=2, cc=1.0, delh=1.0, order=2, periodic=(1, 1, 0))
        sim.set_boundary_condition(2, 0, capi.BC_WALL)
        sim.set_boundary_condition(2, 1, capi.BC_INFLOW, [0, 0, 0, 0.0, 0.1, 0.0])
        problems.setup_uniform_plasma(sim, nd, cd, SHOCK_SPECIES, (32, 32), B0=(0.0, 0.1, 0.0), seed=1)
        return sim, 0.5, "shock 1-D Nx=2^18, 16-cell chunks, 2x32 ppc, wall at x=0 (upstream injection not modelled)"
    raise SystemExit(f"unknown workload {name}")


def main():
    names = sys.argv[1:] or ["thermal3d", "twostream", "cherenkov", "mrx", "shock"]
    peak = 6445.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    stream = torch.cuda.Stream()
    for name in names:
        sim, dt, label = build(name)
        sim.set_stream(stream.cuda_stream)
        npart = int(sim.get_np_all().sum())
        ncell = int(np.prod([s - 2 * sim.nb if s > 1 + 2 * sim.nb else 1 for s in sim.shape])) * sim.nchunk
        steps, warmup = 10, 3
        with torch.cuda.stream(stream):
            sim.step(dt, warmup)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            sim.step(dt, steps)
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1) / steps
            # the fused push + deposit alone
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sim.push_bfd(0.5 * dt)
            k0.record(stream)
            sim.push_deposit_fused(dt)
            k1.record(stream)
            k1.synchronize()
            kms = k0.elapsed_time(k1)
        sim.synchronize()
        regrows, late = sim.growth_stats()
        alg = npart * 120 + ncell * 112
        print(json.dumps({"workload": name, "config": label, "particles": npart, "cells": ncell, "ms_per_step": ms,
                          "particle_steps_per_s": npart / ms * 1e3, "push_deposit_ms": kms,
                          "push_deposit_share": kms / ms, "push_deposit_hbm_frac": alg / (kms * 1e-3) / 1e9 / peak,
                          "particles_after": int(sim.get_np_all().sum()), "segment_regrows": regrows,
                          "late_particles": late}), flush=True)
        sim.close()
        del sim
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
