"""GPU, BASELINE.json's full size (thermal-3D: 128^3 cells in 512 chunks of 16^3, 64 ppc, 1.34e8
particles): size-independent properties of the step, since the reference cannot run this in seconds.

  * particle number conserved in a periodic box, every segment's pindex ends at its Np (sortedness)
  * charge conservation: sum(div E - rho) and sum(div B) stay at round-off (Esirkepov + Yee)
  * total energy (field + particle, from the moments) drifts by < 1e-4 over the steps
  * the fused tiled kernel and the generic thread-per-particle kernels agree on the current
"""
import numpy as np
import pytest

from helpers import cells_consistent
from picnix_b200 import problems

pytestmark = pytest.mark.gpu

CELLS, CHUNK, PPC = 128, 16, (32, 32)


@pytest.fixture(scope="module")
def sim():
    from picnix_b200 import CudaSim

    nd = (CELLS,) * 3
    cd = tuple(n // CHUNK for n in nd)
    s = CudaSim(nd, cd, Ns=2, cc=10.0, delh=1.0, order=2, pusher=0, interp=0)
    problems.setup_uniform_plasma(s, nd, cd, problems.THERMAL_SPECIES, PPC, B0=(5.0, 0.0, 0.0), seed=1)
    yield s
    s.close()


def total_energy(s):
    s.deposit_moment()
    s.exchange(2)
    return s.get_energy().sum()


def test_full_size_conservation(sim):
    n0 = int(sim.get_np_all().sum())
    assert n0 == CELLS ** 3 * sum(PPC)
    e0 = total_energy(sim)
    sim.step(0.05, 5)
    sim.synchronize()
    np_all = sim.get_np_all()
    assert int(np_all.sum()) == n0
    for ic in (0, 137, sim.nchunk - 1):
        for isp in range(2):
            assert sim.get_pindex(ic, isp)[-1] == np_all[ic, isp]
    # every slot lies in the pindex range of the key recomputed from its position
    assert cells_consistent(sim, (CELLS,) * 3, (CELLS // CHUNK,) * 3, chunks=(0, 137, sim.nchunk - 1))
    de = sim.get_diverror()
    scale = np.abs(sim.get_field(0, 1)[..., 0]).max() * CHUNK ** 3  # |rho| summed over a chunk
    assert np.abs(de[:, 0]).max() < 1e-9 * max(scale, 1.0)
    assert np.abs(de[:, 1]).max() < 1e-9
    e1 = total_energy(sim)
    assert abs(e1 - e0) / e0 < 1e-4


def test_full_size_tiled_equals_generic_current(sim):
    """One more fused push+deposit with the tiled kernel; then the same deposit recomputed by the
    generic kernels from (xv, xu) must give the same current."""
    sim.push_bfd(0.025)
    sim.push_velocity(0.05)
    sim.push_position(0.05)
    sim.deposit_current(0.05)          # tiled deposit-only kernel (pindex still valid)
    ja = [sim.get_field(ic, 1) for ic in (0, 255, 511)]
    sim.set_option("force_generic", 1)
    sim.deposit_current(0.05)          # generic kernel, fp64 atomics
    sim.set_option("force_generic", 0)
    jb = [sim.get_field(ic, 1) for ic in (0, 255, 511)]
    for a, b in zip(ja, jb):
        assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b))
