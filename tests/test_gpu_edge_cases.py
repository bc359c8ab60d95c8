"""GPU parity on the edge cases of the path (against the compiled reference):
ragged and empty particle segments, particles that move more than one cell in a step (the
far-mover list of the tiled kernel), every pusher / interpolation variant of the tiled fused
kernel over several steps at relativistic temperature, and a step on a state without particles."""
import numpy as np
import pytest

from helpers import FIELD_UF, FIELD_UJ, counts_equal, field_err, keys_equal, particle_err
from oracle import ref_backend
from picnix_b200 import problems

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_backend.available(), reason="oracle/_ref not built")]


def make_pair(ndims, cdims, species, ppc, cc, thin=None, order=2, pusher=0, interp=0, B0=(5.0, 0.0, 0.0), seed=5):
    """Reference and CUDA arenas with identical state; `thin(chunk id, species)` -> number of
    particles to keep in that segment (None keeps all)."""
    from picnix_b200 import CudaSim

    kw = dict(Ns=len(species), cc=cc, delh=1.0, order=order, pusher=pusher, interp=interp)
    sims = [ref_backend.RefSim(ndims, cdims, vector_mode=1, **kw), CudaSim(ndims, cdims, **kw)]
    dims = problems.chunk_dims(ndims, cdims)
    for sim in sims:
        _, coord = sim.chunkmap()
        for isp, (q, m) in enumerate(problems.species_charge_mass(species, ppc)):
            sim.set_species(isp, q, m)
        nb = sim.nb
        for ic in range(sim.nchunk):
            uf = np.zeros(sim.shape + (6,), dtype=np.float64)
            uf[nb:nb + dims[0], nb:nb + dims[1], nb:nb + dims[2], 3:6] = B0
            sim.set_field(ic, FIELD_UF, uf)
            parts = problems.make_chunk_particles(ic, coord[ic], dims, 1.0, species, ppc, seed)
            for isp, xu in enumerate(parts):
                keep = xu.shape[0] if thin is None else thin(ic, isp, xu.shape[0])
                # capacity as for the full segment, so that arrivals fit
                sim.set_particles(ic, isp, xu[:keep], np_alloc=int(xu.shape[0] * 1.5))
        sim.finalize_setup()
    return sims


def test_ragged_and_empty_segments():
    def thin(ic, isp, n):
        if ic % 3 == 0:
            return 0                    # chunk without any particle
        if isp == 1 and ic % 2 == 1:
            return 7                    # nearly empty species
        return n - (13 * ic) % 50       # ragged
    ref, gpu = make_pair((16, 16, 16), (2, 2, 2), problems.THERMAL_SPECIES, (8, 8), 10.0, thin=thin)
    ref.step(0.05, 12)
    gpu.step(0.05, 12)
    gpu.synchronize()
    assert counts_equal(gpu, ref)
    assert field_err(gpu, ref, FIELD_UF) < 1e-10
    assert field_err(gpu, ref, FIELD_UJ) < 1e-10
    dx, du, same = particle_err(gpu, ref, scale_x=16.0, scale_u=10.0)
    assert same and dx < 1e-11 and du < 1e-11


def test_no_particles_at_all():
    ref, gpu = make_pair((16, 16, 16), (2, 2, 2), problems.THERMAL_SPECIES, (4, 4), 10.0,
                         thin=lambda ic, isp, n: 0)
    ref.step(0.05, 3)
    gpu.step(0.05, 3)
    gpu.synchronize()
    assert counts_equal(gpu, ref)
    assert problems.total_particles(gpu) == 0
    assert field_err(gpu, ref, FIELD_UF) < 1e-13


def test_fused_kernel_far_movers():
    """dt large enough that many particles cross more than one cell: the tiled kernel hands them to
    its far-mover list; current, keys and phase space must still match the reference."""
    ref, gpu = make_pair((16, 16, 16), (2, 2, 2), problems.THERMAL_SPECIES, (8, 8), 10.0)
    dt = 0.9
    ref.push_velocity(dt)
    ref.push_position(dt)
    ref.deposit_current(dt)
    gpu.push_deposit_fused(dt)
    dx, du, same = particle_err(gpu, ref, scale_x=16.0, scale_u=10.0)
    assert same and dx < 1e-14 and du < 1e-13
    assert keys_equal(gpu, ref)
    assert field_err(gpu, ref, FIELD_UJ) < 1e-12


@pytest.mark.parametrize("pusher,interp", [(0, 1), (1, 0), (1, 1), (2, 0), (2, 1)])
def test_tiled_kernel_variants_relativistic_multistep(pusher, interp):
    """Vay / Higuera-Cary pushers and WT interpolation through the tiled fused kernel, at a
    relativistic temperature (u ~ c), over several steps."""
    species = [dict(qm=-1.0, ro=1.0, vt=1.0), dict(qm=+0.1, ro=10.0, vt=0.3)]
    ref, gpu = make_pair((16, 16, 16), (2, 2, 2), species, (8, 8), 1.0, pusher=pusher, interp=interp,
                         B0=(0.5, 0.2, 0.0))
    dt = 0.4  # c dt / dx = 0.4 < 1/sqrt(3)
    ref.step(dt, 8)
    gpu.step(dt, 8)
    gpu.synchronize()
    assert counts_equal(gpu, ref)
    assert field_err(gpu, ref, FIELD_UF) < 1e-10
    assert field_err(gpu, ref, FIELD_UJ) < 1e-10
    dx, du, same = particle_err(gpu, ref, scale_x=16.0, scale_u=1.0)
    assert same and dx < 1e-11 and du < 1e-10


@pytest.mark.parametrize("nspecies", [3, 4, 5])
def test_tiled_3d_kernel_several_species(nspecies):
    """The merged particle stream of the 3-D tiled kernel holds up to four species (the entry table then
    fills all 32 lanes); five fall back to the round-1 kernel.  Drifting beams (the 3-D variant of
    example/beam/twostream) plus extra species, ragged and empty segments, several steps."""
    species = (problems.TWOSTREAM_SPECIES + [dict(qm=+0.5, ro=1.0, vt=0.7, drift=(0.0, 3.0, -2.0)),
                                             dict(qm=-0.2, ro=2.0, vt=0.4, drift=(1.0, 0.0, 4.0))])[:nspecies]
    ppc = (6, 5, 9, 4, 3)[:nspecies]
    thin = lambda ic, isp, n: 0 if (ic == 3 and isp == 1) else (n if (ic + isp) % 3 else n // 2)
    ref, gpu = make_pair((16, 16, 32), (2, 2, 2), species, ppc, 50.0, B0=(10.0, 2.0, 0.0), thin=thin)
    dt = 0.01
    ref.step(dt, 10)
    gpu.step(dt, 10)
    gpu.synchronize()
    assert counts_equal(gpu, ref)
    assert field_err(gpu, ref, FIELD_UF) < 1e-10
    assert field_err(gpu, ref, FIELD_UJ) < 1e-10
    dx, du, same = particle_err(gpu, ref, scale_x=32.0, scale_u=50.0)
    assert same and dx < 1e-11 and du < 1e-10


@pytest.mark.parametrize("ndims,cdims", [((12, 24, 48), (2, 2, 2)),     # chunk 6 x 12 x 24: 3 x-segments, 3 y-groups
                                         ((1, 40, 80), (1, 2, 2)),      # chunk 20 x 40: 5 x-segments, 5 y-groups
                                         ((1, 1, 72), (1, 1, 3)),       # chunk 24: 3 segments per chunk, 9 in all
                                         ((10, 12, 24), (2, 2, 2)),     # chunk x = 12: not a multiple of 8 -> generic
                                         ((1, 18, 32), (1, 3, 2))])     # chunk y = 6: not a multiple of 4 -> generic
def test_chunk_shapes_of_the_tiled_kernels(ndims, cdims):
    """Row geometry of the tiled kernels away from the cubic benchmark chunk: several x-segments and y-groups
    per chunk in every dimensionality, and shapes that do not split (thread-per-particle path)."""
    ref, gpu = make_pair(ndims, cdims, problems.THERMAL_SPECIES, (6, 6), 10.0, B0=(5.0, 1.0, 0.5))
    scale = float(max(ndims))
    ref.step(0.05, 8)
    gpu.step(0.05, 8)
    gpu.synchronize()
    assert counts_equal(gpu, ref)
    assert field_err(gpu, ref, FIELD_UF) < 1e-10
    assert field_err(gpu, ref, FIELD_UJ) < 1e-10
    dx, du, same = particle_err(gpu, ref, scale_x=scale, scale_u=10.0)
    assert same and dx < 1e-11 and du < 1e-10


def test_tiled_2d_kernel_far_movers_and_phases():
    """The 2-D tiled kernel (rowpush2d.cu): one fused push + deposit against the reference's separate
    calls, with a time step so large that many particles cross more than one cell (far-mover list)."""
    ref, gpu = make_pair((1, 32, 32), (1, 2, 2), problems.THERMAL_SPECIES, (8, 8), 10.0)
    dt = 0.9
    ref.push_velocity(dt)
    ref.push_position(dt)
    ref.deposit_current(dt)
    gpu.push_deposit_fused(dt)
    dx, du, same = particle_err(gpu, ref, scale_x=32.0, scale_u=10.0)
    assert same and dx < 1e-14 and du < 1e-13
    assert keys_equal(gpu, ref)
    assert field_err(gpu, ref, FIELD_UJ) < 1e-12
    # the tiled deposit-only path: the reference's separate calls on a fresh pair (old position in xv)
    ref, gpu = make_pair((1, 32, 32), (1, 2, 2), problems.THERMAL_SPECIES, (8, 8), 10.0)
    for sim in (ref, gpu):
        sim.push_velocity(dt)
        sim.push_position(dt)
        sim.deposit_current(dt)
    assert field_err(gpu, ref, FIELD_UJ) < 1e-12


@pytest.mark.parametrize("pusher,interp", [(0, 0), (0, 1), (1, 0), (2, 1)])
def test_tiled_2d_kernel_variants_relativistic_multistep(pusher, interp):
    """Pushers and WT interpolation through the tiled 2-D kernel at a relativistic temperature, three
    species (the merged particle stream holds up to four), ragged segments, several steps."""
    species = [dict(qm=-1.0, ro=1.0, vt=1.0), dict(qm=+0.1, ro=10.0, vt=0.3), dict(qm=-0.5, ro=0.5, vt=0.6)]
    ref, gpu = make_pair((1, 32, 48), (1, 2, 3), species, (8, 5, 3), 1.0, pusher=pusher, interp=interp,
                         B0=(0.5, 0.2, 0.3), thin=lambda ic, isp, n: n if (ic + isp) % 3 else n // 3)
    dt = 0.4
    ref.step(dt, 8)
    gpu.step(dt, 8)
    gpu.synchronize()
    assert counts_equal(gpu, ref)
    assert field_err(gpu, ref, FIELD_UF) < 1e-10
    assert field_err(gpu, ref, FIELD_UJ) < 1e-10
    dx, du, same = particle_err(gpu, ref, scale_x=48.0, scale_u=1.0)
    assert same and dx < 1e-11 and du < 1e-10


def test_tiled_1d_kernel_far_movers_and_phases():
    """The 1-D tiled kernel (rowpush1d.cu): one fused push + deposit against the reference's separate
    calls with a time step so large that many particles cross more than one cell; then the deposit-only
    path through the separate phases.  8-cell chunks: every chunk is one segment, a block spans four."""
    for cdims in ((1, 1, 6), (1, 1, 3)):   # 8- and 16-cell chunks; 6 segments: the last block is ragged
        ref, gpu = make_pair((1, 1, 48), cdims, problems.TWOSTREAM_SPECIES, (16, 16, 32), 50.0, B0=(10.0, 0, 0))
        dt = 0.15   # the beams (u = 10) move 1.5 cells per step
        ref.push_velocity(dt)
        ref.push_position(dt)
        ref.deposit_current(dt)
        gpu.push_deposit_fused(dt)
        dx, du, same = particle_err(gpu, ref, scale_x=48.0, scale_u=50.0)
        assert same and dx < 1e-14 and du < 1e-13
        assert keys_equal(gpu, ref)
        assert field_err(gpu, ref, FIELD_UJ) < 1e-12
        ref, gpu = make_pair((1, 1, 48), cdims, problems.TWOSTREAM_SPECIES, (16, 16, 32), 50.0, B0=(10.0, 0, 0))
        for sim in (ref, gpu):
            sim.push_velocity(dt)
            sim.push_position(dt)
            sim.deposit_current(dt)
        assert field_err(gpu, ref, FIELD_UJ) < 1e-12


@pytest.mark.parametrize("pusher,interp", [(0, 0), (0, 1), (1, 0), (2, 1)])
def test_tiled_1d_kernel_variants_relativistic_multistep(pusher, interp):
    """Pushers and WT interpolation through the tiled 1-D kernel at a relativistic temperature, four
    species (the merged stream's maximum), ragged and empty segments, several steps."""
    species = [dict(qm=-1.0, ro=1.0, vt=1.0), dict(qm=+0.1, ro=10.0, vt=0.3), dict(qm=-0.5, ro=0.5, vt=0.6),
               dict(qm=+0.3, ro=2.0, vt=0.2, drift=(0.5, 0.1, -0.2))]
    thin = lambda ic, isp, n: 0 if (ic == 2 and isp == 1) else (n if (ic + isp) % 3 else n // 3)
    ref, gpu = make_pair((1, 1, 80), (1, 1, 5), species, (8, 5, 3, 6), 1.0, pusher=pusher, interp=interp,
                         B0=(0.5, 0.2, 0.3), thin=thin)
    dt = 0.4
    ref.step(dt, 10)
    gpu.step(dt, 10)
    gpu.synchronize()
    assert counts_equal(gpu, ref)
    assert field_err(gpu, ref, FIELD_UF) < 1e-10
    assert field_err(gpu, ref, FIELD_UJ) < 1e-10
    dx, du, same = particle_err(gpu, ref, scale_x=80.0, scale_u=1.0)
    assert same and dx < 1e-11 and du < 1e-10


def make_cherenkov_pair(ndims, cdims, nppc=16, u0=0.1, vt=0.1, delh=0.1, cc=1.0, order=2, seed=9):
    """example/cherenkov (main.cpp:31-160, config.toml): pair plasma (mime = 1) drifting with u0 along
    x, cell size delh = 0.1 (NOT 1), cc = 1, wp = 1: me = 1/nppc, qe = -wp/nppc*sqrt(gamma)."""
    from picnix_b200 import CudaSim

    gamma = np.sqrt(1 + u0 * u0 / (cc * cc))
    me = 1.0 / nppc
    qe = -1.0 / nppc * np.sqrt(gamma)
    kw = dict(Ns=2, cc=cc, delh=delh, order=order, pusher=0, interp=0)
    sims = [ref_backend.RefSim(ndims, cdims, vector_mode=1, **kw), CudaSim(ndims, cdims, **kw)]
    dims = problems.chunk_dims(ndims, cdims)
    species = [dict(qm=1.0, ro=1.0, vt=vt, drift=(u0, 0.0, 0.0)), dict(qm=1.0, ro=1.0, vt=vt, drift=(u0, 0.0, 0.0))]
    for sim in sims:
        _, coord = sim.chunkmap()
        sim.set_species(0, qe, me)
        sim.set_species(1, -qe, me)
        for ic in range(sim.nchunk):
            sim.set_field(ic, FIELD_UF, np.zeros(sim.shape + (6,), dtype=np.float64))
            parts = problems.make_chunk_particles(ic, coord[ic], dims, delh, species, (nppc, nppc), seed)
            for isp, xu in enumerate(parts):
                sim.set_particles(ic, isp, xu)
        sim.finalize_setup()
    return sims


@pytest.mark.parametrize("ndims,cdims", [((1, 32, 32), (1, 2, 2)), ((16, 16, 16), (2, 2, 2))])
def test_cherenkov_like_small_cells(ndims, cdims):
    """BASELINE configs[2]: cell size 0.1, c = 1, drifting pair plasma; 2-D (generic kernels) and the
    same in 3-D (tiled kernel) -- catches any place that assumes unit cells."""
    ref, gpu = make_cherenkov_pair(ndims, cdims)
    dt = 0.05
    ref.step(dt, 20)
    gpu.step(dt, 20)
    gpu.synchronize()
    assert counts_equal(gpu, ref)
    assert field_err(gpu, ref, FIELD_UF) < 1e-10
    assert field_err(gpu, ref, FIELD_UJ) < 1e-10
    L = 0.1 * max(ndims)
    dx, du, same = particle_err(gpu, ref, scale_x=L, scale_u=1.0)
    assert same and dx < 1e-11 and du < 1e-10
    de_ref, de_gpu = ref.get_diverror().sum(0), gpu.get_diverror().sum(0)
    assert abs(de_gpu[0]) < 1e-9 and abs(de_ref[0]) < 1e-9


def test_lazy_sort_equals_physical_sort():
    """Option lazy_sort (index-only counting sort consumed by the next tiled push) vs the sort that
    moves the particles: same pindex / Np, same particles, same fields after several steps, and the
    pending permutation is materialised transparently by every other entry point."""
    from test_gpu_vs_reference import make_pair as mp

    _, lazy = mp("t3d", perturb=None)
    _, eager = mp("t3d", perturb=None)
    eager.set_option("lazy_sort", 0)
    for sim in (lazy, eager):
        sim.step(0.05, 7)
    # entry points that must see physically ordered arrays: moments, generic deposit, downloads
    for sim in (lazy, eager):
        sim.deposit_moment()
        sim.exchange(2)
    assert field_err(lazy, eager, 3) < 1e-12
    assert counts_equal(lazy, eager)
    dx, du, same = particle_err(lazy, eager, scale_x=16.0, scale_u=10.0)
    assert same and dx < 1e-12 and du < 1e-12
    for sim in (lazy, eager):
        sim.step(0.05, 3)
        sim.synchronize()
    assert counts_equal(lazy, eager)
    assert field_err(lazy, eager, FIELD_UF) < 1e-11
    assert field_err(lazy, eager, FIELD_UJ) < 1e-11
