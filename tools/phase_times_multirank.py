#!/usr/bin/env python
"""Per-phase device time of the overlapped multi-rank step (distributed.step_phases) on every rank:

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/phase_times_multirank.py

CUDA events between the calls of one step, averaged over the timed steps, rank 0's numbers printed
(and the slowest rank's step).  The thermal3d benchmark box, 128^3 cells per GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from picnix_b200.distributed import MODE_CUR, MODE_EMF, MODE_PARTICLE, DistributedSim  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


class A:
    cells = ppc = ref_cells = parity_cells = 0


wl = bench.Workload(sys.argv[1] if len(sys.argv) > 1 else "thermal3d", A)
ndims, cdims = wl.box(wl.dims, world)
sim = DistributedSim(ndims, cdims, rank=rank, world=world, **wl.sim_kwargs())
stream = torch.cuda.Stream()
sim.set_stream(stream.cuda_stream)
wl.setup(sim, ndims, cdims, seed=1, chunk_id_begin=sim.chunk_id_begin)
tr, dt = sim.transport, wl.delt
names, acc = [], {}


def mark(name, evs):
    e = torch.cuda.Event(enable_timing=True)
    e.record(stream)
    evs.append((name, e))


def one_step(evs):
    mark("start", evs)
    sim.push_bfd(0.5 * dt); mark("push_bfd", evs)
    sim.push_deposit_fused(dt); mark("push_deposit", evs)
    sim.boundary_begin(MODE_CUR); t_cur = tr.start(MODE_CUR); mark("begin+start J", evs)
    sim.boundary_begin(MODE_PARTICLE); t_par = tr.start(MODE_PARTICLE); mark("begin+start particles", evs)
    sim.push_bfd(0.5 * dt); mark("push_bfd 2", evs)
    tr.finish(t_cur); mark("wait J", evs)
    sim.boundary_end(MODE_CUR); mark("end J", evs)
    sim.push_efd(dt); mark("push_efd", evs)
    sim.boundary_begin(MODE_EMF); t_emf = tr.start(MODE_EMF); mark("begin+start E/B", evs)
    tr.finish(t_par); mark("wait particles", evs)
    sim.boundary_end(MODE_PARTICLE); mark("end particles (+sort)", evs)
    tr.finish(t_emf); mark("wait E/B", evs)
    sim.boundary_end(MODE_EMF); mark("end E/B", evs)


with torch.cuda.stream(stream):
    sim.commit()
    for _ in range(5):
        one_step([])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    nstep = 20
    for _ in range(nstep):
        evs = []
        one_step(evs)
        torch.cuda.synchronize()
        for (n0, e0), (n1, e1) in zip(evs[:-1], evs[1:]):
            acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1)
            if n1 not in names:
                names.append(n1)
    # where does the HOST spend its time when nothing synchronises?  (a host that blocks inside the step
    # cannot run ahead of the device, and every launch latency after the big kernel becomes visible)
    import time
    cpu = {}

    def cmark(name, evs):
        now = time.perf_counter()
        if evs:
            cpu[name] = cpu.get(name, 0.0) + (now - evs[-1])
        evs.append(now)

    gmark, mark = mark, cmark
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(nstep):
        one_step([])
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    mark = gmark
total = sum(acc.values()) / nstep
t = torch.tensor([total], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"# {wl.name}, {world} rank(s); rank 0 phases, ms (synchronised after every step)")
    for n in names:
        print(f"{n:28s} {acc[n] / nstep:8.3f}")
    print(f"{'step (rank 0)':28s} {total:8.3f}")
    print(f"{'step (slowest rank)':28s} {float(t[0]):8.3f}")
    print(f"# host time per call, ms (free-running, {nstep} steps): enqueue {1e3 * t_enq / nstep:.3f} ms/step, "
          f"wall {1e3 * t_all / nstep:.3f} ms/step")
    for n in names:
        print(f"host {n:23s} {1e3 * cpu.get(n, 0.0) / nstep:8.3f}")
if world > 1:
    dist.destroy_process_group()
