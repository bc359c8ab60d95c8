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


def harris_parameters(cc=1.0, sigma=0.0625, mime=25.0, tite=5.0, lcs=2.5, ncs=50, nbg=10):
    """Derived quantities of the Harris current sheet of example/mrx/main.cpp:21-42."""
    b0 = np.sqrt(sigma)
    qe, qi = -1.0 / ncs, +1.0 / ncs
    me = abs(qe)
    mi = me * mime
    vdi = -cc * b0 / (qi * ncs * lcs) / (1 + tite) * tite
    vde = +cc * b0 / (qi * ncs * lcs) / (1 + tite) * 1.0
    vti = np.sqrt(0.5 * b0 * b0 / (ncs * mi) / (1 + tite) * tite)
    vte = np.sqrt(0.5 * b0 * b0 / (ncs * me) / (1 + tite) * 1.0)
    return dict(b0=b0, q=(qe, qi), m=(me, mi), vd=(vde, vdi), vt=(vte, vti))


def setup_harris_sheet(sim, ndims, cdims, delh=0.2, lcs=2.5, ncs=50, nbg=10, sigma=0.0625, mime=25.0, tite=5.0,
                       bg=0.0, db=0.1, seed=0, chunk_id_begin=0, finalize=True):
    """2-D Harris current sheet between conducting walls in y: the initial condition of example/mrx
    (main.cpp:17-181; cc = 1, two species, phi = 0), with numpy's generator instead of std::mt19937_64.
    The density is far from uniform -- chunks inside the sheet hold (nbg + ncs tanh-profile) particles per
    cell, the others nbg -- which is what the load balancer is for.  The arena must have been created with
    periodic = (1, 0, 1) and PICNIX_BC_CONDUCTING on both y faces."""
    hp = harris_parameters(1.0, sigma, mime, tite, lcs, ncs, nbg)
    dims = chunk_dims(ndims, cdims)
    _, coord = sim.chunkmap()
    for isp in range(2):
        sim.set_species(isp, hp["q"][isp], hp["m"][isp])
    nb = sim.nb
    Mz, My, Mx = sim.shape
    xcs, ycs = 0.5 * ndims[2] * delh, 0.5 * ndims[1] * delh
    numcell = dims[0] * dims[1] * dims[2]
    for ic in range(sim.nchunk):
        gid = chunk_id_begin + ic
        cx, cy, cz = (int(v) for v in coord[gid])
        x0, y0, z0 = cx * dims[2] * delh, cy * dims[1] * delh, cz * dims[0] * delh
        # fields on the whole padded array (main.cpp:56-84)
        xi = x0 + (np.arange(Mx) - nb + 0.5) * delh - xcs
        yi = y0 + (np.arange(My) - nb + 0.5) * delh - ycs
        X, Y = np.meshgrid(xi, yi, indexing="xy")          # [My, Mx]
        az = lambda xx, yy: 2 * db * lcs * np.exp(-(xx * xx + yy * yy) / (4 * lcs * lcs))
        dbx = hp["b0"] * (+(az(X, Y) - az(X, Y - delh)) / delh)
        dby = hp["b0"] * (-(az(X, Y) - az(X - delh, Y)) / delh)
        uf = np.zeros((Mz, My, Mx, 6))
        uf[..., 3] = dbx + hp["b0"] * np.tanh(Y / lcs)
        uf[..., 4] = dby
        uf[..., 5] = hp["b0"] * bg
        sim.set_field(ic, FIELD_UF, uf)
        # particles (main.cpp:101-176): background + current-sheet population, same positions for both species
        rng = np.random.default_rng(seed * 1000003 + gid)
        ymin, ymax = (y0 - ycs) / lcs, (y0 + dims[1] * delh - ycs) / lcs
        rbg = numcell * nbg
        rcs = numcell * ncs * (np.tanh(ymax) - np.tanh(ymin)) / (ymax - ymin)
        mp = int(rbg + rcs)
        x = rng.random(mp) * dims[2] * delh + x0
        z = rng.random(mp) * dims[0] * delh + z0
        sheet = rng.random(mp) < rcs / (rcs + rbg)
        r = rng.random(mp)
        y = np.where(sheet, ycs + lcs * np.arctanh(np.tanh(ymin) + r * (np.tanh(ymax) - np.tanh(ymin))),
                     rng.random(mp) * dims[1] * delh + y0)
        y = np.clip(y, y0, np.nextafter(y0 + dims[1] * delh, y0))
        for isp in range(2):
            xu = np.zeros((mp, 7))
            xu[:, 0], xu[:, 1], xu[:, 2] = x, y, z
            xu[:, 3:6] = rng.normal(size=(mp, 3)) * hp["vt"][isp]
            xu[:, 5] += np.where(sheet, hp["vd"][isp], 0.0)
            xu[:, 6] = (np.int64(mp) * 4 * gid + np.int64(mp) * isp + np.arange(mp, dtype=np.int64)).view(np.float64)
            sim.set_particles(ic, isp, xu)
    if finalize:
        sim.finalize_setup()


def total_particles(sim):
    return int(sum(sim.get_np(ic, isp) for ic in range(sim.nchunk) for isp in range(sim.Ns)))
