"""GPU parity: CUDA path (through the C ABI) against the compiled reference (oracle/_ref).

The reference library is prebuilt in the container where /root/reference exists and travels to the
GPU box as a binary; tests skip themselves when it is absent (the port oracle tests in
test_gpu_vs_port.py do not depend on it).

Tolerances: the reference's own scalar and vector paths differ by ~5e-15 on these
quantities, so
  Maxwell                 1e-15 relative to max |field|
  velocity / position     1e-13 relative to the particle scale
  current                 1e-12 relative to max |J|   (summation order over <= 64*125 terms)
  N-step fields           1e-10 relative to max |field|
integer results (Np, pindex, keys) must be bit-exact.
"""
import numpy as np
import pytest

from helpers import (FIELD_FF, FIELD_UF, FIELD_UJ, MODE_CUR, MODE_EMF, MODE_PARTICLE, cells_consistent,
                     counts_equal, field_err, keys_equal, particle_err)
from oracle import ref_backend
from picnix_b200 import problems

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_backend.available(), reason="oracle/_ref not built")]

CASES = {
    # name: (ndims, cdims, species, ppc, cc, (Bx, By, Bz))
    "t3d": ((16, 16, 16), (2, 2, 2), problems.THERMAL_SPECIES, (8, 8), 10.0, (5.0, 0.0, 0.0)),
    "t2d": ((1, 32, 32), (1, 2, 4), problems.THERMAL_SPECIES, (16, 16), 10.0, (5.0, 0.0, 0.0)),
    "ts1d": ((1, 1, 64), (1, 1, 8), problems.TWOSTREAM_SPECIES, (16, 16, 32), 50.0, (10.0, 0.0, 0.0)),
    # the benchmark's chunk shape and density (bench.py): 16^3-cell chunks, 2 x 32 ppc -- both x segments of
    # a row, four y groups per z, several full batches per row segment and species in the tiled kernel
    "t3d_bench": ((32, 32, 32), (2, 2, 2), problems.THERMAL_SPECIES, (32, 32), 10.0, (5.0, 0.0, 0.0)),
}


def make_pair(case, order=2, pusher=0, interp=0, friedman=0.0, perturb=0.01, seed=3, vector_mode=1,
              periodic=(1, 1, 1)):
    from picnix_b200 import CudaSim

    ndims, cdims, species, ppc, cc, _ = CASES[case]
    kw = dict(Ns=len(species), cc=cc, delh=1.0, order=order, pusher=pusher, interp=interp, friedman=friedman,
              periodic=periodic)
    ref = ref_backend.RefSim(ndims, cdims, vector_mode=vector_mode, **kw)
    gpu = CudaSim(ndims, cdims, **kw)
    for sim in (ref, gpu):
        problems.setup_uniform_plasma(sim, ndims, cdims, species, ppc, B0=CASES[case][5], seed=seed,
                                      perturb=perturb)
    return ref, gpu


@pytest.mark.parametrize("case", ["t3d", "t2d", "ts1d"])
@pytest.mark.parametrize("friedman", [0.0, 0.1])
def test_maxwell(case, friedman):
    ref, gpu = make_pair(case, friedman=friedman)
    dt = 0.05
    # a current to drive E: deposit once on both sides (same particles)
    for sim in (ref, gpu):
        sim.push_velocity(dt)
        sim.push_position(dt)
        sim.deposit_current(dt)
        sim.exchange(MODE_CUR)
    for _ in range(3):
        for sim in (ref, gpu):
            sim.push_bfd(0.5 * dt)
            sim.push_bfd(0.5 * dt)
            sim.push_efd(dt)
            sim.exchange(MODE_EMF)
    assert field_err(gpu, ref, FIELD_UF) < 1e-13
    assert field_err(gpu, ref, FIELD_FF) < 1e-13


@pytest.mark.parametrize("case", ["t3d", "t2d", "ts1d"])
@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_velocity_position_orders(case, order):
    ref, gpu = make_pair(case, order=order)
    dt = 0.05
    for sim in (ref, gpu):
        sim.push_velocity(dt)
        sim.push_position(dt)
    dx, du, same = particle_err(gpu, ref, scale_x=16.0, scale_u=10.0)
    assert same
    assert dx < 1e-14 and du < 1e-13
    # xv holds the state before the position push
    dx, du, same = particle_err(gpu, ref, scale_x=16.0, scale_u=10.0, which=1)
    assert same and dx < 1e-14 and du < 1e-13
    # keys computed by the position push
    assert keys_equal(gpu, ref)


@pytest.mark.parametrize("pusher", [0, 1, 2])
@pytest.mark.parametrize("interp", [0, 1])
def test_velocity_pushers(pusher, interp):
    ref, gpu = make_pair("t3d", pusher=pusher, interp=interp)
    for sim in (ref, gpu):
        sim.push_velocity(0.05)
    dx, du, same = particle_err(gpu, ref, scale_x=16.0, scale_u=10.0)
    assert same and dx == 0.0 and du < 1e-13


@pytest.mark.parametrize("case", ["t3d", "t2d", "ts1d"])
@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_deposit_orders(case, order):
    ref, gpu = make_pair(case, order=order)
    dt = 0.05
    for sim in (ref, gpu):
        sim.push_velocity(dt)
        sim.push_position(dt)
        sim.deposit_current(dt)
    assert field_err(gpu, ref, FIELD_UJ) < 1e-12
    for sim in (ref, gpu):
        sim.exchange(MODE_CUR)
    assert field_err(gpu, ref, FIELD_UJ) < 1e-12


@pytest.mark.parametrize("case", ["t3d", "t2d", "ts1d"])
def test_fused_equals_separate(case):
    ref, gpu = make_pair(case)
    dt = 0.05
    ref.push_velocity(dt)
    ref.push_position(dt)
    ref.deposit_current(dt)
    gpu.push_deposit_fused(dt)
    dx, du, same = particle_err(gpu, ref, scale_x=16.0, scale_u=10.0)
    assert same and dx < 1e-14 and du < 1e-13
    assert field_err(gpu, ref, FIELD_UJ) < 1e-12
    assert keys_equal(gpu, ref)


@pytest.mark.parametrize("case", ["t3d", "t2d", "ts1d"])
def test_migration_and_sort(case):
    ref, gpu = make_pair(case)
    dt = 0.2  # large step: many particles change chunk
    for sim in (ref, gpu):
        sim.push_velocity(dt)
        sim.push_position(dt)
        sim.exchange(MODE_PARTICLE)
    assert counts_equal(gpu, ref)
    dx, du, same = particle_err(gpu, ref, scale_x=16.0, scale_u=10.0)
    assert same and dx < 1e-14 and du < 1e-13


@pytest.mark.parametrize("case,nstep", [("t3d", 20), ("t2d", 20), ("ts1d", 40)])
def test_multistep(case, nstep):
    ref, gpu = make_pair(case, perturb=None)
    dt = 0.05 if case != "ts1d" else 0.01
    ref.step(dt, nstep)
    gpu.step(dt, nstep)
    gpu.synchronize()
    assert counts_equal(gpu, ref)
    assert field_err(gpu, ref, FIELD_UF) < 1e-10
    assert field_err(gpu, ref, FIELD_UJ) < 1e-10
    dx, du, same = particle_err(gpu, ref, scale_x=16.0, scale_u=10.0)
    assert same and dx < 1e-11 and du < 1e-11
    # conservation residuals: sum(div E - rho) and sum(div B) at round-off on both sides
    de_ref, de_gpu = ref.get_diverror().sum(0), gpu.get_diverror().sum(0)
    assert abs(de_gpu[0]) < 1e-10 and abs(de_ref[0]) < 1e-10
    assert abs(de_gpu[1]) < 1e-10 and abs(de_ref[1]) < 1e-10


def test_benchmark_shape_phase_by_phase():
    """16^3-cell chunks at 2 x 32 ppc (the shape bench.py times): one fused tiled push + deposit on
    cell-sorted particles against the reference's three separate calls -- velocity, position, keys,
    current -- then the J halo, the migration and the sort."""
    ref, gpu = make_pair("t3d_bench")
    dt = 0.05
    ref.push_velocity(dt)
    ref.push_position(dt)
    ref.deposit_current(dt)
    gpu.push_deposit_fused(dt)
    dx, du, same = particle_err(gpu, ref, scale_x=32.0, scale_u=10.0)
    assert same and dx < 1e-14 and du < 1e-13
    assert keys_equal(gpu, ref)
    assert field_err(gpu, ref, FIELD_UJ) < 1e-12
    for sim in (ref, gpu):
        sim.exchange(MODE_CUR)
        sim.exchange(MODE_PARTICLE)
    assert field_err(gpu, ref, FIELD_UJ) < 1e-12
    assert counts_equal(gpu, ref)
    assert cells_consistent(gpu, CASES["t3d_bench"][0], CASES["t3d_bench"][1])


@pytest.mark.parametrize("lazy", [1, 0])
def test_benchmark_shape_multistep(lazy):
    """20 whole steps at the benchmark's chunk shape through picnix_cuda_step (fused tiled kernel;
    with the lazy sort the reordering rides on the next push) against PicApplication's schedule run
    by the reference: Np / pindex bit-exact, fields <= 1e-10, phase space matched by id <= 1e-11."""
    ref, gpu = make_pair("t3d_bench", perturb=None)
    gpu.set_option("lazy_sort", lazy)
    ref.step(0.05, 20)
    gpu.step(0.05, 20)
    gpu.synchronize()
    assert counts_equal(gpu, ref)
    assert field_err(gpu, ref, FIELD_UF) < 1e-10
    assert field_err(gpu, ref, FIELD_UJ) < 1e-10
    dx, du, same = particle_err(gpu, ref, scale_x=32.0, scale_u=10.0)
    assert same and dx < 1e-11 and du < 1e-11
    assert cells_consistent(gpu, CASES["t3d_bench"][0], CASES["t3d_bench"][1])
    de_ref, de_gpu = ref.get_diverror().sum(0), gpu.get_diverror().sum(0)
    assert abs(de_gpu[0]) < 1e-9 and abs(de_ref[0]) < 1e-9
    assert abs(de_gpu[1]) < 1e-10 and abs(de_ref[1]) < 1e-10


def test_open_boundary_drops_particles():
    ref, gpu = make_pair("t3d", periodic=(1, 1, 0), perturb=None)
    ref.step(0.05, 10)
    gpu.step(0.05, 10)
    assert counts_equal(gpu, ref)
    n0 = 2 * 8 * 16 ** 3
    assert problems.total_particles(gpu) < n0


def test_mma_deposit_variant_matches_reference():
    """The FP64-MMA formulation of the deposit (option deposit_mma; slower, kept as evidence) passes
    the same parity bar as the scalar row kernel."""
    ref, gpu = make_pair("t3d", perturb=None)
    gpu.set_option("deposit_mma", 1)
    ref.step(0.05, 10)
    gpu.step(0.05, 10)
    gpu.synchronize()
    assert counts_equal(gpu, ref)
    assert field_err(gpu, ref, FIELD_UF) < 1e-10
    assert field_err(gpu, ref, FIELD_UJ) < 1e-10
