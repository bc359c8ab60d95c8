"""GPU: the CUDA path (through the C ABI) against the golden vectors recorded from the compiled reference.

Needs nothing but the repository: no /root/reference, no oracle/_ref.
Tolerances as in test_gpu_vs_reference.py; integer results (keys, pindex, Np) bit-exact.
"""
import pytest

import golden_check as gc

pytestmark = pytest.mark.gpu


def make_cuda(ndims, cdims, **kw):
    from picnix_b200 import CudaSim

    return CudaSim(ndims, cdims, **kw)


@pytest.mark.parametrize("name", gc.golden_cases())
@pytest.mark.parametrize("fused", [False, True])
def test_cuda_phases(name, fused):
    gc.check_phases(make_cuda, name, fused=fused)


@pytest.mark.parametrize("name", gc.golden_cases())
def test_cuda_multistep(name):
    import numpy as np

    sim, g = gc.check_multistep(make_cuda, name)
    # PicChunk::get_energy after deposit_moment, as recorded from the reference (no BoundaryMom
    # exchange in the fixture): field and particle energies per chunk
    sim.deposit_moment()
    assert np.allclose(sim.get_energy(), g["end_energy"], rtol=1e-10, atol=1e-12 * np.abs(g["end_energy"]).max())
