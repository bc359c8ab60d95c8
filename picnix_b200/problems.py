"""Synthetic workloads of BASELINE.json for any backend with the chunk-level API.

The initial conditions follow the reference's example problems:
  * thermal  -- uniform Maxwellian plasma in a uniform B field (example/thermal/main.cpp:17-112)
  * beam     -- the same plus a per-species drift (example/beam/main.cpp:17-118), e.g. the
                two-stream configuration example/beam/twostream/config.toml
Like the reference with seed_type='fixed' every chunk seeds its own generator with its chunk id,
all species of a chunk share the particle positions (charge neutrality), m = ro/np, q = qm*m and
the 64-bit particle id is stored bitwise in component 6.  The generator here is numpy's (the
reference uses std::mt19937_64); parity tests feed the SAME arrays to both sides, so only the
statistics need to match.
"""
import numpy as np

FIELD_UF = 0


def chunk_dims(ndims, cdims):
    return tuple(int(n) // int(c) for n, c in zip(ndims, cdims))


def species_charge_mass(species, ppc):
    """species: list of dict(qm, ro[, vt, drift]); ppc: particles per cell of each species."""
    out = []
    for sp, n in zip(species, ppc):
        m = sp["ro"] / n
        out.append((sp["qm"] * m, m))
    return out


THERMAL_SPECIES = [
    dict(qm=-1.0, ro=1.0, vt=1.0, drift=(0.0, 0.0, 0.0)),
    dict(qm=+0.1, ro=10.0, vt=0.31622776601, drift=(0.0, 0.0, 0.0)),
]

TWOSTREAM_SPECIES = [
    dict(qm=-1.0, ro=0.5, vt=1.0, drift=(10.0, 0.0, 0.0)),
    dict(qm=-1.0, ro=0.5, vt=1.0, drift=(-10.0, 0.0, 0.0)),
    dict(qm=+0.01, ro=100.0, vt=1.0, drift=(0.0, 0.0, 0.0)),
]


def make_chunk_particles(chunk_id, coord_xyz, dims, delh, species, ppc, seed=0):
    """Particles of one chunk: list over species of AoS arrays [np][7]."""
    rng = np.random.default_rng(seed * 1000003 + chunk_id)
    ncell = dims[0] * dims[1] * dims[2]
    cx, cy, cz = (int(v) for v in coord_xyz)
    out = []
    pos_cache = {}
    for isp, (sp, n) in enumerate(zip(species, ppc)):
        mp = n * ncell
        if mp not in pos_cache:
            # same stream for every species -> identical positions (charge neutrality)
            prng = np.random.default_rng(seed * 1000003 + chunk_id)
            pos_cache[mp] = prng.random((mp, 3))
        pos = pos_cache[mp]
        xu = np.zeros((mp, 7), dtype=np.float64)
        xu[:, 0] = pos[:, 0] * (dims[2] * delh) + cx * dims[2] * delh
        xu[:, 1] = pos[:, 1] * (dims[1] * delh) + cy * dims[1] * delh
        xu[:, 2] = pos[:, 2] * (dims[0] * delh) + cz * dims[0] * delh
        vel = rng.normal(size=(mp, 3)) * sp["vt"]
        xu[:, 3:6] = vel + np.asarray(sp.get("drift", (0, 0, 0)), dtype=np.float64)
        ids = np.int64(mp) * chunk_id + np.arange(mp, dtype=np.int64)
        xu[:, 6] = ids.view(np.float64)
        out.append(xu)
    return out


def setup_uniform_plasma(sim, ndims, cdims, species, ppc, delh=1.0, E0=(0, 0, 0), B0=(0, 0, 0), seed=0,
                         chunk_id_begin=0, finalize=True, perturb=None):
    """Fill every local chunk of `sim` with a uniform field and a uniform drifting Maxwellian."""
    dims = chunk_dims(ndims, cdims)
    _, coord = sim.chunkmap()
    for isp, (q, m) in enumerate(species_charge_mass(species, ppc)):
        sim.set_species(isp, q, m)
    nb = sim.nb
    Mz, My, Mx = sim.shape
    for ic in range(sim.nchunk):
        gid = chunk_id_begin + ic
        uf = np.zeros(sim.shape + (6,), dtype=np.float64)
        zs = slice(nb, nb + dims[0])
        ys = slice(nb, nb + dims[1])
        xs = slice(nb, nb + dims[2])
        for k, v in enumerate(tuple(E0) + tuple(B0)):
            uf[zs, ys, xs, k] = v
        if perturb is not None:
            prng = np.random.default_rng(seed * 7919 + 17 * gid + 1)
            uf[zs, ys, xs, :] += perturb * prng.standard_normal(uf[zs, ys, xs, :].shape)
        sim.set_field(ic, FIELD_UF, uf)
        parts = make_chunk_particles(gid, coord[gid], dims, delh, species, ppc, seed)
        for isp, xu in enumerate(parts):
            sim.set_particles(ic, isp, xu)
    if finalize:
        sim.finalize_setup()


def total_particles(sim):
    return int(sum(sim.get_np(ic, isp) for ic in range(sim.nchunk) for isp in range(sim.Ns)))
