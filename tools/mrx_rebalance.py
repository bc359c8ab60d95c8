#!/usr/bin/env python
"""BASELINE.json config 4 (example/mrx): 2-D Harris-sheet reconnection between conducting walls, chunks
load-balanced across the GPUs of one node.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/mrx_rebalance.py

Every rank owns a contiguous range of space-filling-curve chunk ids (one arena per GPU, halos and particle
migration over NCCL).  The run starts from the EVEN split (the same number of chunks per rank,
Balancer::assign_initial with unit loads), measures the step, then rebalances with the reference's own
balancer logic (picnix_assign_initial / picnix_assign_rebalance on the per-chunk particle counts,
nix/balancer.cpp:8-124), moves the chunks between the GPUs (picnix_cuda_chunk_pack/unpack over NCCL) and
measures again.  One JSON line: particles per rank and step time before / after, conservation checks.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from picnix_b200 import capi, problems  # noqa: E402
from picnix_b200.distributed import DistributedSim  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=1024, help="Nx = Ny (the reference's config.toml has 256)")
ap.add_argument("--chunk", type=int, default=16)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--ncs", type=int, default=50)
ap.add_argument("--nbg", type=int, default=10)
args = ap.parse_args()

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

ndims = (1, args.cells, args.cells)
cdims = (1, args.cells // args.chunk, args.cells // args.chunk)
nchunk = cdims[1] * cdims[2]
delt, delh = 0.1, 0.2
kw = dict(Ns=2, cc=1.0, delh=delh, order=2, periodic=(1, 0, 1))


def make(boundary):
    sim = DistributedSim(ndims, cdims, rank=rank, world=world, boundary=boundary, **kw)
    sim.set_stream(torch.cuda.current_stream().cuda_stream)
    for side in (0, 1):
        sim.set_boundary_condition(1, side, capi.BC_CONDUCTING)
    return sim


def allsum(x):
    t = torch.as_tensor(np.asarray(x, dtype=np.float64), device="cuda")
    if world > 1:
        dist.all_reduce(t)
    return t.cpu().numpy()


def timed_steps(sim, n):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        sim.step_phases(delt)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


def per_rank_particles(sim):
    v = np.zeros(world)
    v[rank] = float(sim.get_np_all().sum())
    return allsum(v)


def chunk_loads(sim):
    loads = np.zeros(nchunk)
    loads[sim.chunk_id_begin:sim.chunk_id_begin + sim.nchunk] = sim.get_np_all().sum(axis=1)
    return allsum(loads)


even = capi.assign_initial(np.ones(nchunk), world)
sim = make(even)
problems.setup_harris_sheet(sim, ndims, cdims, delh=delh, ncs=args.ncs, nbg=args.nbg, seed=3,
                            chunk_id_begin=sim.chunk_id_begin)
n0 = per_rank_particles(sim)
timed_steps(sim, 3)                      # warm-up
ms_even = timed_steps(sim, args.steps)
n_even = per_rank_particles(sim)

# Application::rebalance: loads = particles per chunk (PicChunk::reset_load counts Np / Ng, pic_chunk.cpp:124-133)
loads = chunk_loads(sim) + 1.0           # + the cell load of every chunk
balanced = capi.assign_initial(loads, world)
sim = sim.rebalanced(balanced)
timed_steps(sim, 3)
ms_bal = timed_steps(sim, args.steps)
n_bal = per_rank_particles(sim)
sim.synchronize()
de = sim.get_diverror()
div = allsum([float(np.abs(de[:, 0]).sum()), float(np.abs(de[:, 1]).sum())])

if rank == 0:
    imb = lambda v: float(v.max() / v.mean())
    print(json.dumps({
        "problem": f"mrx Harris sheet {args.cells}^2 cells, {nchunk} chunks of {args.chunk}^2, ncs={args.ncs} nbg={args.nbg}",
        "gpus": world, "particles": float(n0.sum()),
        "particles_conserved": bool(n0.sum() == n_even.sum() == n_bal.sum()),
        "boundary_even": [int(b) for b in even], "boundary_balanced": [int(b) for b in balanced],
        "particles_per_rank_even": n_even.tolist(), "particles_per_rank_balanced": n_bal.tolist(),
        "imbalance_max_over_mean": {"even": imb(n_even), "balanced": imb(n_bal)},
        "ms_per_step": {"even": ms_even, "balanced": ms_bal},
        "sum_abs_divE_minus_rho": float(div[0]), "sum_abs_divB": float(div[1])}))
if world > 1:
    dist.destroy_process_group()
