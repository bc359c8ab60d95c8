"""GPU parity of SURVEY §8(f) N1: PicChunk::deposit_moment, the BoundaryMom exchange and
PicChunk::get_energy against the compiled reference.

Tolerances: um differs from the reference by summation order only (<= 64 particles x 27 points per
cell): 1e-12 of max |um|; energies 1e-12 relative.  The energy-conservation residual of the north
star (total = field + particle energy over N steps) must match the reference's to 1e-10 relative.
"""
import numpy as np
import pytest

from helpers import MODE_MOM, field_err
from oracle import ref_backend
from test_gpu_vs_reference import make_pair

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_backend.available(), reason="oracle/_ref not built")]

FIELD_UM = 3


@pytest.mark.parametrize("case", ["t3d", "t2d", "ts1d"])
@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_deposit_moment_orders(case, order):
    ref, gpu = make_pair(case, order=order)
    for sim in (ref, gpu):
        sim.deposit_moment()
    assert field_err(gpu, ref, FIELD_UM) < 1e-12
    for sim in (ref, gpu):
        sim.exchange(MODE_MOM)
    assert field_err(gpu, ref, FIELD_UM) < 1e-12
    e_ref, e_gpu = ref.get_energy(), gpu.get_energy()
    assert np.allclose(e_gpu, e_ref, rtol=1e-12, atol=1e-12 * np.abs(e_ref).max())


def test_deposit_moment_generic_equals_cell_kernel():
    """The thread-per-particle fallback and the warp-per-cell kernel give the same moments."""
    _, a = make_pair("t3d")
    _, b = make_pair("t3d")
    b.set_option("force_generic", 1)
    for sim in (a, b):
        sim.deposit_moment()
        sim.exchange(MODE_MOM)
    assert field_err(a, b, FIELD_UM) < 1e-12


@pytest.mark.parametrize("case,dt,nstep", [("t3d", 0.05, 20), ("ts1d", 0.01, 40)])
def test_energy_history_matches_reference(case, dt, nstep):
    """history diagnostic: field + particle energy every 5 steps, GPU vs reference."""
    ref, gpu = make_pair(case, perturb=None)
    hist = {"ref": [], "gpu": []}
    for k in range(nstep // 5):
        for name, sim in (("ref", ref), ("gpu", gpu)):
            sim.step(dt, 5)
            sim.deposit_moment()
            sim.exchange(MODE_MOM)
            hist[name].append(sim.get_energy().sum(axis=0))
    h_ref, h_gpu = np.array(hist["ref"]), np.array(hist["gpu"])
    tot_ref, tot_gpu = h_ref.sum(axis=1), h_gpu.sum(axis=1)
    assert np.allclose(h_gpu, h_ref, rtol=1e-10, atol=1e-10 * np.abs(h_ref).max())
    # the residual itself: relative drift of the total energy is the same on both sides
    drift_ref = (tot_ref - tot_ref[0]) / tot_ref[0]
    drift_gpu = (tot_gpu - tot_gpu[0]) / tot_gpu[0]
    assert np.max(np.abs(drift_gpu - drift_ref)) < 1e-10
