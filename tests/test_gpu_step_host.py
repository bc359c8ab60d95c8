"""GPU parity of the host-buffer entry point picnix_cuda_step_host (pipelined H2D / step / D2H):
host arrays in the reference's layouts go in, the same arrays come back advanced by one step, and
must match the compiled reference advanced by the same steps (same tolerances as the resident
multistep test), for pinned and for pageable (page-locked on first use) host buffers."""
import numpy as np
import pytest

from helpers import FIELD_FF, FIELD_UF, FIELD_UJ, sorted_by_id
from oracle import ref_backend
from test_gpu_vs_reference import make_pair

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_backend.available(), reason="oracle/_ref not built")]


def host_field(gpu, st, ic, which):
    ncell = gpu.Ng
    name, width = {FIELD_UF: ("uf", 6), FIELD_UJ: ("uj", 4), FIELD_FF: ("ff", 18)}[which]
    return st[name][ic * ncell * width:(ic + 1) * ncell * width].reshape(gpu._field_shape(which))


@pytest.mark.parametrize("case,dt,nstep", [("t3d", 0.05, 6), ("t2d", 0.05, 6), ("ts1d", 0.01, 10)])
@pytest.mark.parametrize("pinned", [True, False])
def test_step_host_matches_reference(case, dt, nstep, pinned):
    ref, gpu = make_pair(case, perturb=None, friedman=0.05)
    st = gpu.host_state(pinned=pinned)
    for _ in range(nstep):
        gpu.step_host(st, dt, 1)   # full state crosses the boundary every step
    ref.step(dt, nstep)
    for ic in range(gpu.nchunk):
        for which, tol in ((FIELD_UF, 1e-10), (FIELD_UJ, 1e-10), (FIELD_FF, 1e-10)):
            a, b = host_field(gpu, st, ic, which), ref.get_field(ic, which)
            if which == FIELD_FF:
                a, b = a[..., :3], b[..., :3]
            assert np.max(np.abs(a - b)) <= tol * max(np.max(np.abs(b)), 1e-300), (ic, which)
        for isp in range(gpu.Ns):
            assert st["np"][ic * gpu.Ns + isp] == ref.get_np(ic, isp)
            pa = sorted_by_id(gpu.host_particles(st, gpu.Ns, ic, isp))
            pb = sorted_by_id(ref.get_particles(ic, isp))
            assert np.array_equal(pa[:, 6].view(np.int64), pb[:, 6].view(np.int64))
            assert np.max(np.abs(pa[:, :3] - pb[:, :3])) < 1e-11 * 16.0
            assert np.max(np.abs(pa[:, 3:6] - pb[:, 3:6])) < 1e-11 * 10.0


def test_step_host_multi_step_call_equals_resident():
    """nstep > 1 inside one call == the resident picnix_cuda_step on a second arena."""
    _, gpu = make_pair("t3d", perturb=None)
    _, res = make_pair("t3d", perturb=None)
    st = gpu.host_state(pinned=True)
    gpu.step_host(st, 0.05, 5)
    res.step(0.05, 5)
    res.synchronize()
    for ic in range(gpu.nchunk):
        a, b = host_field(gpu, st, ic, FIELD_UF), res.get_field(ic, FIELD_UF)
        assert np.max(np.abs(a - b)) <= 1e-11 * np.max(np.abs(b))
        for isp in range(gpu.Ns):
            assert st["np"][ic * gpu.Ns + isp] == res.get_np(ic, isp)
            assert np.array_equal(gpu.get_pindex(ic, isp), res.get_pindex(ic, isp))


def test_step_host_many_small_segments_grouped_transfers():
    """Arenas with thousands of segments (1-D runs) move their particles in groups of consecutive segments
    (one host-span copy + one transposition launch per group, hostio.cu): a two-stream box of 768 8-cell
    chunks x 3 species = 2304 segments through picnix_cuda_step_host equals the resident step."""
    from picnix_b200 import CudaSim, problems

    nd, cd = (1, 1, 8 * 768), (1, 1, 768)
    sims = []
    for _ in range(2):
        sim = CudaSim(nd, cd, Ns=3, cc=50.0, delh=1.0, order=2)
        problems.setup_uniform_plasma(sim, nd, cd, problems.TWOSTREAM_SPECIES, (16, 16, 32), B0=(10.0, 0, 0), seed=3)
        sims.append(sim)
    gpu, res = sims
    st = gpu.host_state(pinned=True)
    for _ in range(4):
        gpu.step_host(st, 0.01, 1)
    res.step(0.01, 4)
    res.synchronize()
    for ic in range(0, gpu.nchunk, 37):
        a, b = host_field(gpu, st, ic, FIELD_UF), res.get_field(ic, FIELD_UF)
        assert np.max(np.abs(a - b)) <= 1e-11 * max(np.max(np.abs(b)), 1e-300)
        for isp in range(gpu.Ns):
            assert st["np"][ic * gpu.Ns + isp] == res.get_np(ic, isp)
            pa = sorted_by_id(gpu.host_particles(st, gpu.Ns, ic, isp))
            pb = sorted_by_id(res.get_particles(ic, isp))
            assert np.array_equal(pa[:, 6].view(np.int64), pb[:, 6].view(np.int64))
            assert np.max(np.abs(pa[:, :6] - pb[:, :6])) < 1e-9
    assert np.array_equal(st["np"].reshape(-1, 3), res.get_np_all())


def test_host_alloc_roundtrip():
    import ctypes as C

    from picnix_b200 import capi

    lib = capi.load()
    p = C.c_void_p()
    assert lib.picnix_cuda_host_alloc(C.byref(p), 1 << 20) == capi.OK and p.value
    assert lib.picnix_cuda_host_free(p) == capi.OK


def test_upload_download_state_roundtrip():
    """picnix_cuda_upload_state / download_state (the snapshot path): a state downloaded from one
    arena and uploaded into a fresh one continues to the same result."""
    _, a = make_pair("t3d", perturb=None)
    _, b = make_pair("t3d", perturb=None, seed=11)   # different initial particles: must be overwritten
    a.step(0.05, 3)
    st = a.host_state(pinned=True)
    a.download_state(st)
    b.upload_state(st)
    a.step(0.05, 3)
    b.step(0.05, 3)
    a.synchronize()
    b.synchronize()
    for ic in range(a.nchunk):
        fa, fb = a.get_field(ic, FIELD_UF), b.get_field(ic, FIELD_UF)
        assert np.max(np.abs(fa - fb)) <= 1e-11 * np.max(np.abs(fa))
        for isp in range(a.Ns):
            assert a.get_np(ic, isp) == b.get_np(ic, isp)
            assert np.array_equal(a.get_pindex(ic, isp), b.get_pindex(ic, isp))
