#!/usr/bin/env python
"""Two (or more) GPUs: load rebalancing over NCCL (DistributedSim.rebalanced) checked against a
single-arena run of the same non-uniform problem that every rank also computes on its own GPU.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/rebalance_demo.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from picnix_b200 import CudaSim, capi, problems  # noqa: E402
from picnix_b200.distributed import DistributedSim, MODE_EMF  # noqa: E402
from test_gpu_multirank_one_device import fill  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))

ndims, cdims = (16, 32, 16), (2, 4, 2)
species, ppc, B0, dt = problems.THERMAL_SPECIES, (8, 8), (5.0, 0.0, 0.0), 0.05
kw = dict(Ns=2, cc=10.0, delh=1.0, order=2)

single = CudaSim(ndims, cdims, **kw)
fill(single, ndims, cdims, species, ppc, B0)
single.exchange(MODE_EMF)

even = capi.assign_initial(np.ones(single.nchunk), world)
sim = DistributedSim(ndims, cdims, rank=rank, world=world, boundary=even, **kw)
fill(sim, ndims, cdims, species, ppc, B0)
sim.exchange(MODE_EMF)
sim.step(dt, 5)

loads = torch.zeros(single.nchunk, dtype=torch.float64, device="cuda")
loads[sim.chunk_id_begin:sim.chunk_id_begin + sim.nchunk] = torch.from_numpy(
    sim.get_np_all().sum(axis=1).astype(np.float64)).cuda()
dist.all_reduce(loads)
balanced = capi.assign_initial(loads.cpu().numpy(), world)
before = float(sim.get_np_all().sum())
sim = sim.rebalanced(balanced)
after = float(sim.get_np_all().sum())
sim.step(dt, 5)
sim.synchronize()

single.step(dt, 10)
single.synchronize()
worst = 0.0
ok = True
for ic in range(sim.nchunk):
    gid = sim.chunk_id_begin + ic
    for which in (0, 1):
        a, b = sim.get_field(ic, which), single.get_field(gid, which)
        worst = max(worst, float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)))
    for isp in range(2):
        ok = ok and sim.get_np(ic, isp) == single.get_np(gid, isp)
        ok = ok and np.array_equal(sim.get_pindex(ic, isp), single.get_pindex(gid, isp))
ok = ok and worst < 1e-11
res = torch.tensor([float(ok), worst, before, after], dtype=torch.float64, device="cuda")
allres = [torch.zeros_like(res) for _ in range(world)]
dist.all_gather(allres, res)
if rank == 0:
    print(json.dumps({"even": even.tolist(), "balanced": balanced.tolist(),
                      "ok": all(bool(r[0]) for r in allres), "max_rel_field_err": max(float(r[1]) for r in allres),
                      "particles_per_rank_before": [float(r[2]) for r in allres],
                      "particles_per_rank_after": [float(r[3]) for r in allres]}))
dist.destroy_process_group()
