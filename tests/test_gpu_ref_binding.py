"""The drop-in, actually dropped in: the reference's OWN application
(nix::Application::main -> PicApplication::push_openmp -> example/thermal MainChunk, config.toml
driven) is run twice with the same configuration file

  * host/ref_binding/_build/thermal_ref   -- unmodified, PicChunk on the host CPU
  * host/ref_binding/_build/thermal_cuda  -- the same main.cpp with MainChunk deriving from
    CudaPicChunk (host/ref_binding/cuda_pic_chunk.hpp), i.e. every kernel of the step on the B200
    through the C ABI, history / field / particle diagnostics written by the reference's own writers

and the outputs are compared: history.txt (div errors, field and particle energies; the file holds 7
significant digits), the raw field dumps (uf and the moments um, full precision) and the raw particle
dumps (matched by particle id).  Both binaries are built by host/ref_binding/Makefile in the container
that has the reference tree and travel to the GPU box; the test skips when they are absent.
"""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "host", "ref_binding", "_build")
REF, CUDA = os.path.join(BUILD, "thermal_ref"), os.path.join(BUILD, "thermal_cuda")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(CUDA)),
                                 reason="host/ref_binding/_build is not built (needs the reference tree)")]

CONFIG = """
[application]
  basedir = 'data'
  [application.log]
    interval = 100
  [application.rebalance]
    interval = 1000000
  [application.option]
    vectorization = 'vector'
    seed_type = 'fixed'
    order = {order}

[[diagnostic]]
  name = 'history'
  interval = 1

[[diagnostic]]
  name = 'field'
  interval = {nstep}

[[diagnostic]]
  name = 'particle'
  interval = {nstep}
  fraction = 1.0

[parameter]
  Nx = {nx}
  Ny = {ny}
  Nz = {nz}
  Cx = {cx}
  Cy = {cy}
  Cz = {cz}
  Ex = 0.0
  Ey = 0.0
  Ez = 0.0
  Bx = 5.0
  By = 0.0
  Bz = 0.0
  Ns = 2
  cc = 10.0
  delt = 0.05
  delh = 1.0

[[parameter.particle]]
    np = {ppc}
    qm = -1.0
    ro = 1.0
    vt = 1.0

[[parameter.particle]]
    np = {ppc}
    qm = +0.1
    ro = 10.0
    vt = 0.31622776601
"""


def run_app(binary, workdir, cfg, nstep):
    os.makedirs(workdir, exist_ok=True)
    with open(os.path.join(workdir, "config.toml"), "w") as fp:
        fp.write(cfg)
    env = dict(os.environ, OMP_NUM_THREADS="4", PICNIX_SYNC_HOST_INTERVAL="1")
    proc = subprocess.run([binary, "-c", "config.toml", "-t", str(0.05 * nstep)], cwd=workdir, env=env,
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-3000:]
    return os.path.join(workdir, "data")


def read_history(path):
    rows = [line.split() for line in open(os.path.join(path, "history.txt")) if not line.startswith("#")]
    return np.array([[float(v) for v in r] for r in rows])


def read_dump(path, kind, step):
    meta = json.load(open(os.path.join(path, kind, f"{step:08d}.json")))
    raw = np.fromfile(os.path.join(path, kind, meta["meta"]["rawfile"]), dtype=np.uint8)
    out = {}
    for name, ds in meta["dataset"].items():
        if ds["datatype"] != "f8":
            continue
        out[name] = raw[ds["offset"]:ds["offset"] + ds["size"]].view(np.float64).reshape(ds["shape"])
    return out


@pytest.mark.parametrize("shape", ["3d", "2d"])
def test_reference_application_with_cuda_chunks(tmp_path, shape):
    nstep = 20
    geom = dict(nx=16, ny=16, nz=16, cx=2, cy=2, cz=2, ppc=8) if shape == "3d" else \
        dict(nx=32, ny=32, nz=1, cx=4, cy=2, cz=1, ppc=16)
    cfg = CONFIG.format(order=2, nstep=nstep, **geom)
    ref = run_app(REF, str(tmp_path / "ref"), cfg, nstep)
    gpu = run_app(CUDA, str(tmp_path / "gpu"), cfg, nstep)

    # history.txt: step, time, div(E), div(B), E^2/2, B^2/2, particle energies (%13.6e each)
    ha, hb = read_history(gpu), read_history(ref)
    assert ha.shape == hb.shape and ha.shape[0] >= nstep + 1
    assert np.array_equal(ha[:, :2], hb[:, :2])
    assert np.max(np.abs(ha[:, 2:4])) < 1e-10 and np.max(np.abs(hb[:, 2:4])) < 1e-10  # round-off on both sides
    scale = np.maximum(np.abs(hb[:, 4:]), 1e-300)
    assert np.max(np.abs(ha[:, 4:] - hb[:, 4:]) / scale) < 3e-6  # the printed precision

    # full precision: raw field dump of the last step (uf interior and moments um of every chunk)
    fa, fb = read_dump(gpu, "field", nstep), read_dump(ref, "field", nstep)
    assert set(fa) == set(fb) and "uf" in fa
    for name in fa:
        assert fa[name].shape == fb[name].shape
        assert np.max(np.abs(fa[name] - fb[name])) <= 1e-10 * np.max(np.abs(fb[name])), name

    # raw particle dump, matched by the 64-bit id stored in component 6
    pa, pb = read_dump(gpu, "particle", nstep), read_dump(ref, "particle", nstep)
    assert set(pa) == set(pb) and len(pa) >= 2
    for name in pa:
        a, b = pa[name].reshape(-1, 7), pb[name].reshape(-1, 7)
        assert a.shape == b.shape
        a = a[np.argsort(a[:, 6].view(np.int64), kind="stable")]
        b = b[np.argsort(b[:, 6].view(np.int64), kind="stable")]
        assert np.array_equal(a[:, 6].view(np.int64), b[:, 6].view(np.int64))
        assert np.max(np.abs(a[:, :3] - b[:, :3])) < 1e-11 * 32
        assert np.max(np.abs(a[:, 3:6] - b[:, 3:6])) < 1e-11 * 10


MRX_REF, MRX_CUDA = os.path.join(BUILD, "mrx_ref"), os.path.join(BUILD, "mrx_cuda")

MRX_CONFIG = """
[application]
  basedir = 'data'
  [application.log]
    interval = 100
  [application.rebalance]
    interval = 1000000
  [application.option]
    vectorization = 'vector'
    seed_type = 'fixed'
    order = 2

[[diagnostic]]
  name = 'history'
  interval = 1

[[diagnostic]]
  name = 'field'
  interval = {nstep}

[[diagnostic]]
  name = 'particle'
  interval = {nstep}
  fraction = 1.0

[parameter]
  Nx = 64
  Ny = 32
  Nz = 1
  Cx = 4
  Cy = 2
  Cz = 1
  Ns = 2
  delt = 0.1
  delh = 0.2
  lcs = 2.5
  ncs = 8
  nbg = 4
  mime = 25
  sigma = 0.0625
  tite = 5.0
  bg = 0.0
  db = 0.1
  phi = 0.0
"""


@pytest.mark.skipif(not (os.path.exists(MRX_REF) and os.path.exists(MRX_CUDA)),
                    reason="host/ref_binding/_build/mrx_* not built")
def test_mrx_application_with_cuda_chunks(tmp_path):
    """BASELINE configs[3]: the reference's example/mrx application (Harris sheet between conducting walls,
    MainChunk::setup with its own RNG, MainApplication with non-periodic y) unmodified on the CPU and with
    its chunks on the B200 (host/ref_binding/mrx_cuda.cpp: the example's wall hooks replaced by
    PICNIX_BC_CONDUCTING).  history, raw field dump and raw particle dump (as sets: the example assigns no
    particle ids) must agree."""
    nstep = 30
    cfg = MRX_CONFIG.format(nstep=nstep)

    def run(binary, workdir):
        os.makedirs(workdir, exist_ok=True)
        with open(os.path.join(workdir, "config.toml"), "w") as fp:
            fp.write(cfg)
        env = dict(os.environ, OMP_NUM_THREADS="4", PICNIX_SYNC_HOST_INTERVAL="1")
        proc = subprocess.run([binary, "-c", "config.toml", "-t", str(0.1 * nstep)], cwd=workdir, env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert proc.returncode == 0, proc.stdout[-3000:]
        return os.path.join(workdir, "data")

    ref, gpu = run(MRX_REF, str(tmp_path / "ref")), run(MRX_CUDA, str(tmp_path / "gpu"))
    ha, hb = read_history(gpu), read_history(ref)
    assert ha.shape == hb.shape and ha.shape[0] >= nstep + 1
    assert np.array_equal(ha[:, :2], hb[:, :2])
    assert np.max(np.abs(ha[:, 2:4])) < 1e-10 and np.max(np.abs(hb[:, 2:4])) < 1e-10
    scale = np.maximum(np.abs(hb[:, 4:]), 1e-3 * np.abs(hb[:, 4:]).max(axis=0))  # E^2/2 starts at zero
    assert np.max(np.abs(ha[:, 4:] - hb[:, 4:]) / scale) < 1e-5                    # the printed precision

    fa, fb = read_dump(gpu, "field", nstep), read_dump(ref, "field", nstep)
    assert set(fa) == set(fb) and "uf" in fa
    for name in fa:
        assert fa[name].shape == fb[name].shape
        assert np.max(np.abs(fa[name] - fb[name])) <= 1e-10 * np.max(np.abs(fb[name])), name

    pa, pb = read_dump(gpu, "particle", nstep), read_dump(ref, "particle", nstep)
    assert set(pa) == set(pb) and len(pa) >= 2
    for name in pa:
        a, b = pa[name].reshape(-1, 7), pb[name].reshape(-1, 7)
        assert a.shape == b.shape and a.shape[0] > 0
        a = a[np.lexsort(np.round(a[:, :6].T[::-1], 6))]
        b = b[np.lexsort(np.round(b[:, :6].T[::-1], 6))]
        assert np.max(np.abs(a[:, :3] - b[:, :3])) < 1e-10 * 12.8
        assert np.max(np.abs(a[:, 3:6] - b[:, 3:6])) < 1e-10


TWOSTREAM_CONFIG = """
# example/beam/twostream/config.toml as shipped (BASELINE configs[0]); two changes: seed_type = 'fixed'
# (the default seeds every chunk from std::random_device, which no second run can reproduce) and no in-run
# rebalancing (one rank; the binding moves chunks through picnix_cuda_chunk_pack/unpack between runs)
[application]
  basedir = 'data'
  [application.log]
    interval = 5000
  [application.rebalance]
    interval = 1000000
  [application.option]
    vectorization = 'vector'
    seed_type = 'fixed'

[[diagnostic]]
  name = 'history'
  interval = 10

[[diagnostic]]
  name = 'field'
  interval = {nstep}

[[diagnostic]]
  name = 'particle'
  interval = {nstep}
  fraction = 1.0

[parameter]
  Nx = 512
  Ny = 1
  Nz = 1
  Cx = 64
  Cy = 1
  Cz = 1
  Ex = 0.0
  Ey = 0.0
  Ez = 0.0
  Bx = 10.0
  By = 0.0
  Bz = 0.0
  Ns = 3
  cc = 50.0
  delt = 0.01
  delh = 1.0

[[parameter.particle]]
    np = 16
    qm = -1.0
    ro = 0.5
    vt = 1.0
    vx = 10.0
    vy = 0.0
    vz = 0.0

[[parameter.particle]]
    np = 16
    qm = -1.0
    ro = 0.5
    vt = 1.0
    vx = -10.0
    vy = 0.0
    vz = 0.0

[[parameter.particle]]
    np = 32
    qm = +0.01
    ro = 100.0
    vt = 1.0
    vx = 0.0
    vy = 0.0
    vz = 0.0
"""

CHERENKOV_CONFIG = """
# example/cherenkov/config.toml as shipped (BASELINE configs[2]) with seed_type = 'fixed', no in-run
# rebalancing and the diagnostics at the end of this short run
[application]
  basedir = 'data'
  [application.log]
    interval = 200
  [application.option]
    vectorization = 'vector'
    seed_type = 'fixed'
  [application.rebalance]
    interval = 1000000

[[diagnostic]]
  interval = 5
  name = 'history'

[[diagnostic]]
  interval = {nstep}
  name = 'field'

[[diagnostic]]
  interval = {nstep}
  name = 'particle'
  fraction = 1.0

[parameter]
  Cx = 8
  Cy = 8
  Cz = 1
  Ns = 2
  Nx = 128
  Ny = 128
  Nz = 1
  cc = 1.0
  delh = 0.1
  delt = 0.05
  mime = 1
  nppc = 32
  phi = 0.0
  sigma = 0.0
  theta = 0.0
  u0 = 0.1
  vte = 0.1
  vti = 0.1
  wp = 1.0
"""


@pytest.mark.parametrize("app,config,dt,nstep,length", [("beam", TWOSTREAM_CONFIG, 0.01, 200, 512.0),
                                                        ("cherenkov", CHERENKOV_CONFIG, 0.05, 40, 12.8)])
def test_shipped_configurations_through_the_reference_application(tmp_path, app, config, dt, nstep, length):
    """BASELINE configs[0] and [2] with the parameters the reference ships: the reference's application on
    the CPU and the same application with CudaPicChunk on the B200 (1-D and 2-D tiled kernels, lazy sort,
    halos, migration, moments for the history energies) agree on history, raw fields and raw particles."""
    ref_bin, cuda_bin = os.path.join(BUILD, app + "_ref"), os.path.join(BUILD, app + "_cuda")
    if not (os.path.exists(ref_bin) and os.path.exists(cuda_bin)):
        pytest.skip("host/ref_binding/_build/%s_* not built" % app)
    cfg = config.format(nstep=nstep)
    out = {}
    for tag, binary in (("ref", ref_bin), ("gpu", cuda_bin)):
        workdir = str(tmp_path / tag)
        os.makedirs(workdir, exist_ok=True)
        with open(os.path.join(workdir, "config.toml"), "w") as fp:
            fp.write(cfg)
        env = dict(os.environ, OMP_NUM_THREADS="8", PICNIX_SYNC_HOST_INTERVAL="1")
        proc = subprocess.run([binary, "-c", "config.toml", "-t", str(dt * nstep)], cwd=workdir, env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
        assert proc.returncode == 0, proc.stdout[-3000:]
        out[tag] = os.path.join(workdir, "data")
    ha, hb = read_history(out["gpu"]), read_history(out["ref"])
    assert ha.shape == hb.shape and ha.shape[0] >= 3
    assert np.array_equal(ha[:, :2], hb[:, :2])
    scale = np.maximum(np.abs(hb[:, 4:]), 1e-3 * np.abs(hb[:, 4:]).max(axis=0))
    assert np.max(np.abs(ha[:, 4:] - hb[:, 4:]) / scale) < 1e-5      # the printed precision
    # the Gauss residual is whatever the initial placement left; it must be the same number on both sides
    assert np.max(np.abs(ha[:, 2:4] - hb[:, 2:4])) < 1e-9 * max(1.0, np.max(np.abs(hb[:, 2:4])))

    fa, fb = read_dump(out["gpu"], "field", nstep), read_dump(out["ref"], "field", nstep)
    assert set(fa) == set(fb) and "uf" in fa
    for name in fa:
        assert fa[name].shape == fb[name].shape
        assert np.max(np.abs(fa[name] - fb[name])) <= 1e-9 * np.max(np.abs(fb[name])), name

    pa, pb = read_dump(out["gpu"], "particle", nstep), read_dump(out["ref"], "particle", nstep)
    assert set(pa) == set(pb) and len(pa) >= 2
    for name in pa:
        a, b = pa[name].reshape(-1, 7), pb[name].reshape(-1, 7)
        assert a.shape == b.shape and a.shape[0] > 0
        ida, idb = a[:, 6].view(np.int64), b[:, 6].view(np.int64)
        if np.unique(idb).size == idb.size:                       # the example assigns ids: match by id
            a, b = a[np.argsort(ida, kind="stable")], b[np.argsort(idb, kind="stable")]
            assert np.array_equal(a[:, 6].view(np.int64), b[:, 6].view(np.int64))
        else:                                                     # no ids: compare as sets
            a = a[np.lexsort(np.round(a[:, :6].T[::-1], 6))]
            b = b[np.lexsort(np.round(b[:, :6].T[::-1], 6))]
        assert np.max(np.abs(a[:, :3] - b[:, :3])) < 1e-9 * length
        assert np.max(np.abs(a[:, 3:6] - b[:, 3:6])) < 1e-8
