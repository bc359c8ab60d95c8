"""GPU: physical boundary conditions of the non-periodic BASELINE problems against
the reference's OWN problem code.  The oracle's chunks are the MainChunk of example/mrx/main.cpp and of
example/shock/main.cpp, compiled unchanged into oracle/_ref (oracle/ref_driver.cpp), so its
set_boundary_field / set_boundary_particle hooks are the reference's; the device uses the built-in
boundary kinds of picnix_cuda_set_boundary_condition (picnix_b200/csrc/boundary.cu).

  * conducting walls in y with an arbitrary (uniform, perturbed) plasma between them
  * the Harris current sheet built by the mrx example's own setup() (non-uniform density: the chunks
    in the sheet hold three times the particles of the others), copied to the device
  * the shock tube's wall at the lower x boundary and imposed upstream fields at the upper one, with a
    plasma streaming into the wall
Tolerances as in test_gpu_vs_reference.py: counts / pindex bit-exact, fields 1e-10, phase space 1e-11.
"""
import numpy as np
import pytest

from helpers import FIELD_UF, FIELD_UJ, counts_equal, field_err, particle_err
from oracle import ref_backend
from picnix_b200 import capi, problems

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_backend.available(), reason="oracle/_ref not built")]


def copy_state(ref, gpu, Ns):
    """State of the oracle (built by an example's own setup) -> device arena."""
    lib = ref.lib
    import ctypes as C

    for isp in range(Ns):
        q, m = C.c_double(), C.c_double()
        lib.ref_get_species(ref.h, isp, C.byref(q), C.byref(m))
        gpu.set_species(isp, q.value, m.value)
    for ic in range(ref.nchunk):
        gpu.set_field(ic, FIELD_UF, ref.get_field(ic, FIELD_UF))
        for isp in range(Ns):
            gpu.set_particles(ic, isp, ref.get_particles(ic, isp))
    gpu.finalize_setup()


def phase_space_err(a, b, scale_x, scale_u):
    """Particles compared as SETS per (chunk, species): the mrx example gives its particles no id
    (component 6 stays 0), so rows are ordered by their phase-space coordinates instead."""
    dx = du = 0.0
    for ic in range(a.nchunk):
        for isp in range(a.Ns):
            pa, pb = a.get_particles(ic, isp), b.get_particles(ic, isp)
            assert pa.shape == pb.shape
            if pa.shape[0] == 0:
                continue
            ka = np.lexsort(np.round(pa[:, :6].T[::-1], 6))
            kb = np.lexsort(np.round(pb[:, :6].T[::-1], 6))
            pa, pb = pa[ka], pb[kb]
            dx = max(dx, float(np.max(np.abs(pa[:, 0:3] - pb[:, 0:3])) / scale_x))
            du = max(du, float(np.max(np.abs(pa[:, 3:6] - pb[:, 3:6])) / scale_u))
    return dx, du


def compare(gpu, ref, scale_x, scale_u, by_id=True):
    assert counts_equal(gpu, ref)
    assert field_err(gpu, ref, FIELD_UF) < 1e-10
    assert field_err(gpu, ref, FIELD_UJ) < 1e-10
    if by_id:
        dx, du, same = particle_err(gpu, ref, scale_x=scale_x, scale_u=scale_u)
        assert same
    else:
        dx, du = phase_space_err(gpu, ref, scale_x, scale_u)
    assert dx < 1e-11 and du < 1e-11


def test_conducting_walls_uniform_plasma():
    from picnix_b200 import CudaSim

    ndims, cdims = (1, 32, 32), (1, 4, 2)
    kw = dict(Ns=2, cc=10.0, delh=1.0, order=2, pusher=0, interp=0, periodic=(1, 0, 1))
    ref = ref_backend.RefSim(ndims, cdims, vector_mode=1, problem="mrx", **kw)
    gpu = CudaSim(ndims, cdims, **kw)
    for side in (0, 1):
        gpu.set_boundary_condition(1, side, capi.BC_CONDUCTING)
    for sim in (ref, gpu):
        problems.setup_uniform_plasma(sim, ndims, cdims, problems.THERMAL_SPECIES, (16, 16), B0=(5.0, 0.0, 1.0),
                                      seed=4, perturb=0.01)
    n0 = problems.total_particles(gpu)
    ref.step(0.05, 30)
    gpu.step(0.05, 30)
    gpu.synchronize()
    assert problems.total_particles(gpu) == n0        # walls reflect: nobody leaves
    compare(gpu, ref, 32.0, 10.0)


def test_harris_sheet_from_the_example_setup():
    from picnix_b200 import CudaSim

    ndims, cdims = (1, 32, 64), (1, 4, 4)
    pj = dict(example_setup=True, delt=0.1, delh=0.2, lcs=1.0, ncs=24, nbg=6, sigma=0.0625, mime=25.0, tite=5.0,
              bg=0.0, db=0.1, phi=0.0)
    ref = ref_backend.RefSim(ndims, cdims, Ns=2, cc=1.0, delh=0.2, periodic=(1, 0, 1), vector_mode=1,
                             problem="mrx", problem_json=pj)
    gpu = CudaSim(ndims, cdims, Ns=2, cc=1.0, delh=0.2, periodic=(1, 0, 1))
    for side in (0, 1):
        gpu.set_boundary_condition(1, side, capi.BC_CONDUCTING)
    copy_state(ref, gpu, 2)
    ref.lib.ref_finalize_setup(ref.h)
    npc = gpu.get_np_all()
    assert npc.max() > 2 * npc.min()                  # current-sheet chunks are much heavier
    ref.step(0.1, 40)
    gpu.step(0.1, 40)
    gpu.synchronize()
    assert int(gpu.get_np_all().sum()) == int(npc.sum())
    compare(gpu, ref, 64 * 0.2, 1.0, by_id=False)


def test_shock_tube_wall_and_inflow_fields():
    from picnix_b200 import CudaSim

    ndims, cdims = (1, 1, 64), (1, 1, 8)
    up = dict(Ex0=0.0, Ey0=0.02, Ez0=-0.01, Bx0=0.3, By0=0.1, Bz0=0.2)
    bj = dict(boundary=dict(influx=[0, 0], efflux=[0, 0], lastid=0, nppc=0, delt=0.05, u0=0.1, vte=0.1, vti=0.05, **up))
    kw = dict(Ns=2, cc=1.0, delh=1.0, order=2, pusher=0, interp=0, periodic=(1, 1, 0))
    ref = ref_backend.RefSim(ndims, cdims, vector_mode=1, problem="shock", problem_json=bj, **kw)
    gpu = CudaSim(ndims, cdims, **kw)
    gpu.set_boundary_condition(2, 0, capi.BC_WALL)
    gpu.set_boundary_condition(2, 1, capi.BC_INFLOW, [up[k] for k in ("Ex0", "Ey0", "Ez0", "Bx0", "By0", "Bz0")])
    species = [dict(qm=-1.0, ro=1.0), dict(qm=+0.04, ro=25.0)]
    dims = problems.chunk_dims(ndims, cdims)
    for sim in (ref, gpu):
        _, coord = sim.chunkmap()
        for isp, (q, m) in enumerate(problems.species_charge_mass(species, (16, 16))):
            sim.set_species(isp, q, m)
        for ic in range(sim.nchunk):
            uf = np.zeros(sim.shape + (6,))
            uf[..., 3:6] = (up["Bx0"], up["By0"], up["Bz0"])
            sim.set_field(ic, FIELD_UF, uf)
            parts = problems.make_chunk_particles(ic, coord[ic], dims, 1.0,
                                                  [dict(vt=0.05, drift=(-0.4, 0.0, 0.0)),
                                                   dict(vt=0.02, drift=(-0.4, 0.0, 0.0))], (16, 16), seed=9)
            for isp, xu in enumerate(parts):
                if coord[ic][0] >= 6:              # nothing near the upper boundary: its re-injection draws
                    xu = xu[:0]                    # host random numbers (example/shock/main.cpp:436-535)
                sim.set_particles(ic, isp, xu, np_alloc=1024)
        sim.finalize_setup()
    n0 = problems.total_particles(gpu)
    ref.step(0.05, 200)                               # 0.4 c towards the wall: four cells of plasma bounce
    gpu.step(0.05, 200)
    gpu.synchronize()
    assert problems.total_particles(gpu) == n0
    compare(gpu, ref, 64.0, 1.0)
