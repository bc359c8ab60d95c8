"""Shared comparison helpers for the parity tests."""
import numpy as np

from picnix_b200 import problems

FIELD_UF, FIELD_UJ, FIELD_FF = 0, 1, 2
MODE_EMF, MODE_CUR, MODE_MOM, MODE_PARTICLE = 0, 1, 2, 3


def build_pair(make_a, make_b, ndims, cdims, species, ppc, seed=1, **setup_kw):
    """Create two backends with identical configuration and identical initial state."""
    a, b = make_a(), make_b()
    for sim in (a, b):
        problems.setup_uniform_plasma(sim, ndims, cdims, species, ppc, seed=seed, **setup_kw)
    return a, b


def rel_err(x, y):
    """max |x - y| relative to max |y| (1 if y == 0 everywhere)."""
    scale = max(np.max(np.abs(y)), 1e-300)
    return float(np.max(np.abs(x - y)) / scale) if x.size else 0.0


def field_err(a, b, which, sl=None):
    worst = 0.0
    for ic in range(a.nchunk):
        fa, fb = a.get_field(ic, which), b.get_field(ic, which)
        if which == FIELD_FF:
            fa, fb = fa[..., :3], fb[..., :3]
        if sl is not None:
            fa, fb = fa[sl], fb[sl]
        scale = max(np.max(np.abs(fb)), 1e-300)
        worst = max(worst, float(np.max(np.abs(fa - fb)) / scale))
    return worst


def interior(sim):
    nb = sim.nb
    Mz, My, Mx = sim.shape
    def s(M):
        return slice(nb, M - nb) if M > 1 + 2 * nb else slice(nb, nb + 1)
    return (s(Mz), s(My), s(Mx))


def sorted_by_id(xu):
    ids = xu[:, 6].view(np.int64)
    order = np.argsort(ids, kind="stable")
    return xu[order]


def particle_err(a, b, scale_x=1.0, scale_u=1.0, which=0):
    """Compare particle phase space matched by id; returns (max dx, max du, all ids equal)."""
    dx = du = 0.0
    same_ids = True
    for ic in range(a.nchunk):
        for isp in range(a.Ns):
            pa = sorted_by_id(a.get_particles(ic, isp, which))
            pb = sorted_by_id(b.get_particles(ic, isp, which))
            if pa.shape != pb.shape or not np.array_equal(pa[:, 6].view(np.int64), pb[:, 6].view(np.int64)):
                same_ids = False
                continue
            if pa.shape[0] == 0:
                continue
            dx = max(dx, float(np.max(np.abs(pa[:, 0:3] - pb[:, 0:3])) / scale_x))
            du = max(du, float(np.max(np.abs(pa[:, 3:6] - pb[:, 3:6])) / scale_u))
    return dx, du, same_ids


def counts_equal(a, b):
    """Np and pindex bit-exact for every (chunk, species)."""
    for ic in range(a.nchunk):
        for isp in range(a.Ns):
            if a.get_np(ic, isp) != b.get_np(ic, isp):
                return False
            if not np.array_equal(a.get_pindex(ic, isp), b.get_pindex(ic, isp)):
                return False
    return True


def chunk_limits(sim, ndims, cdims, ic, delh=1.0, chunk_id_begin=0):
    """(zmin, ymin, xmin) of local chunk `ic` from the chunk map (Chunk::set_coordinate)."""
    _, coord = sim.chunkmap()
    cx, cy, cz = (int(v) for v in coord[chunk_id_begin + ic])
    dims = problems.chunk_dims(ndims, cdims)
    return cz * dims[0] * delh, cy * dims[1] * delh, cx * dims[2] * delh


def cells_consistent(sim, ndims=None, cdims=None, delh=1.0, chunk_id_begin=0, chunks=None):
    """Every particle sits in the slot range pindex assigns to the key of its position (sortedness):
    the keys are recomputed here from the downloaded positions and the chunk limits
    (XtensorParticle::count, nix/xtensor_particle.hpp:324-357; even orders) and
    pindex[key] <= slot < pindex[key + 1] is asserted for every slot.  Without ndims/cdims only
    pindex[-1] == Np and monotonicity are checked."""
    for ic in (range(sim.nchunk) if chunks is None else chunks):
        for isp in range(sim.Ns):
            n = sim.get_np(ic, isp)
            pindex = sim.get_pindex(ic, isp).astype(np.int64)
            if pindex[-1] != n or pindex[0] != 0 or np.any(np.diff(pindex) < 0):
                return False
            if ndims is None or n == 0:
                continue
            xu = sim.get_particles(ic, isp, 0, n)
            zmin, ymin, xmin = chunk_limits(sim, ndims, cdims, ic, delh, chunk_id_begin)
            dims = problems.chunk_dims(ndims, cdims)
            has = [d > 1 for d in dims]
            # flatindex strides, nix/xtensor_particle.hpp:231-238 (ignorable dimensions have one cell)
            sy = dims[2] + 1 if has[2] else 2
            sz = sy * ((dims[1] + 1) if has[1] else 2)
            ix = np.floor((xu[:, 0] - xmin) / delh).astype(np.int64) if has[2] else np.zeros(n, np.int64)
            iy = np.floor((xu[:, 1] - ymin) / delh).astype(np.int64) if has[1] else np.zeros(n, np.int64)
            iz = np.floor((xu[:, 2] - zmin) / delh).astype(np.int64) if has[0] else np.zeros(n, np.int64)
            key = iz * sz + iy * sy + ix
            slot = np.arange(n, dtype=np.int64)
            if np.any(key < 0) or np.any(key >= pindex.size - 1):
                return False
            if np.any(slot < pindex[key]) or np.any(slot >= pindex[key + 1]):
                return False
    return True


def keys_equal(a, b):
    """Cell keys (gindex) bit-exact, matched by particle id (order inside a cell is free)."""
    for ic in range(a.nchunk):
        for isp in range(a.Ns):
            pa, pb = a.get_particles(ic, isp), b.get_particles(ic, isp)
            ka, kb = a.get_gindex(ic, isp), b.get_gindex(ic, isp)
            if pa.shape != pb.shape:
                return False
            oa = np.argsort(pa[:, 6].view(np.int64), kind="stable")
            ob = np.argsort(pb[:, 6].view(np.int64), kind="stable")
            if not np.array_equal(ka[oa], kb[ob]):
                return False
    return True
