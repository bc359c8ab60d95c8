"""GPU: the multi-rank path of the CUDA library on ONE device.

R arenas (nrank = R, rank = r) live in one process; the per-peer send buffers the library packs in
boundary_begin are copied device-to-device into the peers' receive buffers (what NCCL does between
GPUs in picnix_b200/distributed.py), in the order of PicApplication::push_openmp.  The result must
equal the single-arena run of the same problem: remote halo pack/unpack, migration records, counts.

Covers BASELINE configs[3] in spirit: a Harris-sheet-like NON-UNIFORM density, with the rank
boundaries taken from Balancer::assign_initial on the particle loads (load-balanced chunk ranges).
"""
import numpy as np
import pytest

from picnix_b200 import capi, problems
from picnix_b200.distributed import MODE_CUR, MODE_EMF, MODE_MOM, MODE_PARTICLE, cuda_view

pytestmark = pytest.mark.gpu


def harris_keep(coord_xyz, cdims, n):
    """fraction of the uniform population kept in a chunk: a current sheet in the middle of y"""
    cy = coord_xyz[1]
    y = (cy + 0.5) / cdims[1] - 0.5
    return max(8, int(n * (0.2 + 0.8 / np.cosh(y / 0.15) ** 2)))


def fill(sim, ndims, cdims, species, ppc, B0, seed=3):
    dims = problems.chunk_dims(ndims, cdims)
    _, coord = sim.chunkmap()
    for isp, (q, m) in enumerate(problems.species_charge_mass(species, ppc)):
        sim.set_species(isp, q, m)
    nb = sim.nb
    for ic in range(sim.nchunk):
        gid = sim.chunk_id_begin + ic
        uf = np.zeros(sim.shape + (6,), dtype=np.float64)
        uf[nb:nb + dims[0], nb:nb + dims[1], nb:nb + dims[2], 3:6] = B0
        sim.set_field(ic, 0, uf)
        parts = problems.make_chunk_particles(gid, coord[gid], dims, 1.0, species, ppc, seed)
        for isp, xu in enumerate(parts):
            keep = harris_keep(coord[gid], cdims, xu.shape[0])
            sim.set_particles(ic, isp, xu[:keep], np_alloc=int(xu.shape[0] * 1.6))
    sim.commit()
    sim.init_friedman()
    sim.sort_particle()


def move_all(sims, mode):
    """send buffer of (rank r -> peer p) into p's receive buffer for r"""
    import torch

    for s in sims:
        s.synchronize()
    peers = [s.peers() for s in sims]
    for r, s in enumerate(sims):
        for i, p in enumerate(peers[r]):
            sp, sb, _, _ = s.comm_buffer(mode, i)
            j = peers[p].index(r)
            if mode == MODE_PARTICLE:
                sims[p].set_recv_bytes(mode, j, sb)
            _, _, rp, rb = sims[p].comm_buffer(mode, j)
            assert rb == sb, (mode, r, p, sb, rb)
            if sb > 0:
                cuda_view(rp, rb).copy_(cuda_view(sp, sb))
    torch.cuda.synchronize()


def exchange_all(sims, mode):
    for s in sims:
        s.boundary_begin(mode)
    move_all(sims, mode)
    for s in sims:
        s.boundary_end(mode)


def step_all(sims, dt):
    for s in sims:
        s.push_bfd(0.5 * dt)
        s.push_deposit_fused(dt)
        s.boundary_begin(MODE_CUR)
    move_all(sims, MODE_CUR)
    for s in sims:
        s.boundary_begin(MODE_PARTICLE)
    move_all(sims, MODE_PARTICLE)
    for s in sims:
        s.push_bfd(0.5 * dt)
        s.boundary_end(MODE_CUR)
        s.push_efd(dt)
        s.boundary_begin(MODE_EMF)
    move_all(sims, MODE_EMF)
    for s in sims:
        s.boundary_end(MODE_PARTICLE)
        s.boundary_end(MODE_EMF)


@pytest.mark.parametrize("async_migration", [0, 1])
@pytest.mark.parametrize("ndims,cdims,nrank", [((1, 32, 32), (1, 4, 4), 3), ((16, 16, 16), (2, 2, 2), 2),
                                               ((16, 32, 16), (2, 4, 2), 4), ((1, 1, 128), (1, 1, 16), 3)])
def test_ranks_on_one_device_equal_single_arena(ndims, cdims, nrank, async_migration):
    from picnix_b200 import CudaSim

    species, ppc, cc, B0, dt, nstep = problems.THERMAL_SPECIES, (8, 8), 10.0, (5.0, 0.0, 0.0), 0.05, 12
    kw = dict(Ns=2, cc=cc, delh=1.0, order=2)

    single = CudaSim(ndims, cdims, **kw)
    fill(single, ndims, cdims, species, ppc, B0)
    single.exchange(MODE_EMF)
    loads = single.get_np_all().sum(axis=1).astype(np.float64)
    boundary = capi.assign_initial(loads, nrank)          # load-balanced SFC ranges
    assert boundary[0] == 0 and boundary[-1] == single.nchunk and np.all(np.diff(boundary) > 0)

    sims = [CudaSim(ndims, cdims, nrank=nrank, rank=r, boundary=boundary, **kw) for r in range(nrank)]
    for s in sims:
        assert s.chunk_id_begin == boundary[s.cfg.rank] and s.nchunk == boundary[s.cfg.rank + 1] - boundary[s.cfg.rank]
        s.set_option("async_migration", async_migration)  # lagged-count particle exchange (no host sync)
        fill(s, ndims, cdims, species, ppc, B0)
    exchange_all(sims, MODE_EMF)

    single.step(dt, nstep)
    single.synchronize()
    for _ in range(nstep):
        step_all(sims, dt)

    moved = 0
    for s in sims:
        s.synchronize()
        for ic in range(s.nchunk):
            gid = s.chunk_id_begin + ic
            for which in (0, 1):
                a, b = s.get_field(ic, which), single.get_field(gid, which)
                assert np.max(np.abs(a - b)) <= 1e-11 * max(np.max(np.abs(b)), 1e-300), (gid, which)
            for isp in range(2):
                assert s.get_np(ic, isp) == single.get_np(gid, isp)
                assert np.array_equal(s.get_pindex(ic, isp), single.get_pindex(gid, isp))
                pa, pb = s.get_particles(ic, isp), single.get_particles(gid, isp)
                ia, ib = np.argsort(pa[:, 6].view(np.int64)), np.argsort(pb[:, 6].view(np.int64))
                assert np.array_equal(pa[ia, 6].view(np.int64), pb[ib, 6].view(np.int64))
                assert np.max(np.abs(pa[ia, :6] - pb[ib, :6])) < 1e-10
                moved += int(np.sum((pa[:, 6].view(np.int64) // (pa.shape[0] + 1)) >= 0))
    assert moved > 0

    # moments across the rank boundaries (BoundaryMom: Ns * 14 components per cell)
    single.deposit_moment()
    single.exchange(MODE_MOM)
    for s in sims:
        s.deposit_moment()
    exchange_all(sims, MODE_MOM)
    for s in sims:
        for ic in range(s.nchunk):
            a, b = s.get_field(ic, 3), single.get_field(s.chunk_id_begin + ic, 3)
            assert np.max(np.abs(a - b)) <= 1e-11 * np.max(np.abs(b))


def test_rebalance_moves_chunks_between_arenas():
    """start from an even split of a non-uniform plasma, run, let the balancer move
    the rank boundaries on the particle loads, move the chunks (pack -> device buffer -> unpack),
    continue -- the result equals the single-arena run."""
    from picnix_b200 import CudaSim
    from picnix_b200.distributed import rebalance_in_process

    ndims, cdims, nrank = (1, 32, 32), (1, 4, 4), 3
    species, ppc, B0, dt = problems.THERMAL_SPECIES, (8, 8), (5.0, 0.0, 0.0), 0.05
    kw = dict(Ns=2, cc=10.0, delh=1.0, order=2)
    single = CudaSim(ndims, cdims, **kw)
    fill(single, ndims, cdims, species, ppc, B0)
    single.exchange(MODE_EMF)

    even = capi.assign_initial(np.ones(single.nchunk), nrank)
    sims = [CudaSim(ndims, cdims, nrank=nrank, rank=r, boundary=even, **kw) for r in range(nrank)]
    for s in sims:
        fill(s, ndims, cdims, species, ppc, B0)
    exchange_all(sims, MODE_EMF)
    for _ in range(5):
        step_all(sims, dt)

    loads = np.concatenate([s.get_np_all().sum(axis=1) for s in sims]).astype(np.float64)
    balanced = capi.assign_initial(loads, nrank)
    assert not np.array_equal(balanced, even)            # the sheet makes the even split unbalanced
    assert np.array_equal(capi.assign_rebalance(loads, even).shape, even.shape)
    sims = rebalance_in_process(sims, balanced,
                                lambda r, b: CudaSim(ndims, cdims, nrank=nrank, rank=r, boundary=b, **kw))
    per_rank = [float(s.get_np_all().sum()) for s in sims]
    assert max(per_rank) / (sum(per_rank) / nrank) < 1.35  # balanced on particle count
    for _ in range(5):
        step_all(sims, dt)

    single.step(dt, 10)
    single.synchronize()
    for s in sims:
        s.synchronize()
        for ic in range(s.nchunk):
            gid = s.chunk_id_begin + ic
            for which in (0, 1):
                a, b = s.get_field(ic, which), single.get_field(gid, which)
                assert np.max(np.abs(a - b)) <= 1e-11 * max(np.max(np.abs(b)), 1e-300), (gid, which)
            for isp in range(2):
                assert s.get_np(ic, isp) == single.get_np(gid, isp)
                assert np.array_equal(s.get_pindex(ic, isp), single.get_pindex(gid, isp))


def test_foreign_particle_record_is_an_error():
    """A received migration record whose destination chunk this rank does not own (a decomposition or
    message-plan mismatch between ranks) must not vanish silently: errflag[2] -> PICNIX_ERR_INVALID at the
    next synchronize, and one step late through the per-step statistics for a host that never synchronises."""
    import torch

    from picnix_b200 import CudaSim

    ndims, cdims, nrank = (16, 16, 16), (2, 2, 2), 2
    kw = dict(Ns=2, cc=10.0, delh=1.0, order=2)
    boundary = capi.assign_initial(np.ones(8), nrank)
    sims = [CudaSim(ndims, cdims, nrank=nrank, rank=r, boundary=boundary, **kw) for r in range(nrank)]
    for s in sims:
        s.set_option("async_migration", 0)
        fill(s, ndims, cdims, problems.THERMAL_SPECIES, (8, 8), (5.0, 0.0, 0.0))
    exchange_all(sims, MODE_EMF)
    step_all(sims, 0.05)
    for s in sims:
        s.push_bfd(0.025)
        s.push_deposit_fused(0.05)
        s.boundary_begin(MODE_PARTICLE)
    move_all(sims, MODE_PARTICLE)
    # corrupt the destination chunk id of the first record rank 1 received (int32 at byte 56 of the record)
    _, _, rp, rb = sims[1].comm_buffer(MODE_PARTICLE, 0)
    assert rb >= 64
    cuda_view(rp, rb)[56:60].copy_(torch.tensor(np.array([10 ** 6], dtype=np.int32).view(np.uint8)))
    torch.cuda.synchronize()
    sims[0].boundary_end(MODE_PARTICLE)
    sims[1].boundary_end(MODE_PARTICLE)
    sims[0].synchronize()
    with pytest.raises(Exception, match="does not own"):
        sims[1].synchronize()


def test_conducting_walls_across_ranks_equal_single_arena():
    """Physical boundary kinds in a multi-rank run: the mrx geometry (periodic x, conducting walls in y) split
    over three ranks equals the single-arena run (which tests/test_gpu_boundaries.py pins against the example)."""
    from picnix_b200 import CudaSim

    ndims, cdims, nrank = (1, 32, 32), (1, 4, 4), 3
    species, ppc, B0, dt, nstep = problems.THERMAL_SPECIES, (8, 8), (5.0, 0.0, 1.0), 0.05, 12
    kw = dict(Ns=2, cc=10.0, delh=1.0, order=2, periodic=(1, 0, 1))

    def walls(sim):
        for side in (0, 1):
            sim.set_boundary_condition(1, side, capi.BC_CONDUCTING)

    single = CudaSim(ndims, cdims, **kw)
    walls(single)
    fill(single, ndims, cdims, species, ppc, B0)
    single.exchange(MODE_EMF)
    boundary = capi.assign_initial(single.get_np_all().sum(axis=1).astype(np.float64), nrank)
    sims = [CudaSim(ndims, cdims, nrank=nrank, rank=r, boundary=boundary, **kw) for r in range(nrank)]
    for s in sims:
        walls(s)
        s.set_option("async_migration", 1)
        fill(s, ndims, cdims, species, ppc, B0)
    exchange_all(sims, MODE_EMF)
    n0 = sum(int(s.get_np_all().sum()) for s in sims)
    single.step(dt, nstep)
    single.synchronize()
    for _ in range(nstep):
        step_all(sims, dt)
    assert sum(int(s.get_np_all().sum()) for s in sims) == n0      # the walls reflect: nobody leaves
    for s in sims:
        s.synchronize()
        for ic in range(s.nchunk):
            gid = s.chunk_id_begin + ic
            for which in (0, 1):
                a, b = s.get_field(ic, which), single.get_field(gid, which)
                assert np.max(np.abs(a - b)) <= 1e-11 * max(np.max(np.abs(b)), 1e-300), (gid, which)
            for isp in range(2):
                assert s.get_np(ic, isp) == single.get_np(gid, isp)
                assert np.array_equal(s.get_pindex(ic, isp), single.get_pindex(gid, isp))
