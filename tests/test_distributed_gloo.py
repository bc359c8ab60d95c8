"""World-size-2 CPU test of the multi-rank path (gloo backend, no GPU).

Two processes each own half of the space-filling-curve chunk ids (like two MPI ranks of the
reference, nix/application.cpp:287-292), hold them in the CPU oracle and move the per-peer halo and
migration buffers with `picnix_b200.distributed.Transport` -- the same transport and the same step
schedule (`step_phases`) the GPU run uses with NCCL.  The result must be IDENTICAL (bit for bit,
including particle order) to the single-rank oracle run of the same problem.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = {
    "t3d": ((8, 8, 8), (2, 2, 2), "thermal", (4, 4), 10.0, (5.0, 0.0, 0.0), 0.1, 5),
    "t2d": ((1, 16, 16), (1, 2, 4), "thermal", (4, 4), 10.0, (5.0, 0.0, 0.0), 0.1, 5),
    "ts1d": ((1, 1, 64), (1, 1, 8), "twostream", (8, 8, 16), 50.0, (10.0, 0.0, 0.0), 0.02, 10),
}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(sim, case, chunk_id_begin):
    from picnix_b200 import problems

    ndims, cdims, spname, ppc, cc, B0, dt, nstep = CASES[case]
    species = problems.THERMAL_SPECIES if spname == "thermal" else problems.TWOSTREAM_SPECIES
    problems.setup_uniform_plasma(sim, ndims, cdims, species, ppc, B0=B0, seed=5, perturb=0.01,
                                  chunk_id_begin=chunk_id_begin, finalize=False)


def _worker(rank, world, port, case, outdir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from oracle import port_backend
    from picnix_b200 import distributed, problems

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ndims, cdims, spname, ppc, cc, B0, dt, nstep = CASES[case]
    species = problems.THERMAL_SPECIES if spname == "thermal" else problems.TWOSTREAM_SPECIES
    sim = port_backend.PortSim(ndims, cdims, Ns=len(species), cc=cc, nrank=world, rank=rank, nthread=1)
    tr = distributed.Transport(sim, world, device_buffers=False)
    _setup(sim, case, sim.chunk_id_begin)
    sim.init_friedman()
    sim.sort_particle()
    distributed.exchange(sim, tr, distributed.MODE_EMF)
    for _ in range(nstep):
        distributed.step_phases(sim, tr, dt)
    sim.deposit_moment()
    distributed.exchange(sim, tr, distributed.MODE_MOM)
    out = {"begin": np.array(sim.chunk_id_begin), "nchunk": np.array(sim.nchunk), "peers": np.array(sim.peers())}
    for ic in range(sim.nchunk):
        out[f"uf_{ic}"] = sim.get_field(ic, 0)
        out[f"uj_{ic}"] = sim.get_field(ic, 1)
        out[f"um_{ic}"] = sim.get_field(ic, 3)
        for isp in range(sim.Ns):
            out[f"xu_{ic}_{isp}"] = sim.get_particles(ic, isp)
            out[f"pindex_{ic}_{isp}"] = sim.get_pindex(ic, isp)
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), **out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["t3d", "t2d", "ts1d"])
def test_two_ranks_equal_one_rank(case, tmp_path):
    import torch.multiprocessing as mp

    from oracle import port_backend
    from picnix_b200 import problems

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)

    ndims, cdims, spname, ppc, cc, B0, dt, nstep = CASES[case]
    species = problems.THERMAL_SPECIES if spname == "thermal" else problems.TWOSTREAM_SPECIES
    single = port_backend.PortSim(ndims, cdims, Ns=len(species), cc=cc, nthread=1)
    _setup(single, case, 0)
    single.finalize_setup()
    single.step(dt, nstep)
    single.deposit_moment()
    single.exchange(2)

    seen = 0
    for rank in range(world):
        g = np.load(os.path.join(str(tmp_path), f"rank{rank}.npz"))
        begin, nchunk = int(g["begin"]), int(g["nchunk"])
        assert list(g["peers"]) == [1 - rank]
        for ic in range(nchunk):
            gid = begin + ic
            assert np.array_equal(g[f"uf_{ic}"], single.get_field(gid, 0)), (case, "uf", gid)
            assert np.array_equal(g[f"uj_{ic}"], single.get_field(gid, 1)), (case, "uj", gid)
            assert np.array_equal(g[f"um_{ic}"], single.get_field(gid, 3)), (case, "um", gid)
            for isp in range(single.Ns):
                assert np.array_equal(g[f"pindex_{ic}_{isp}"], single.get_pindex(gid, isp))
                a, b = g[f"xu_{ic}_{isp}"], single.get_particles(gid, isp)
                assert a.shape == b.shape and np.array_equal(a.view(np.int64), b.view(np.int64)), (case, "xu", gid)
            seen += 1
    assert seen == single.nchunk
