"""Replay a golden fixture (tests/golden/*.npz, produced from the compiled reference) on a backend."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def config_of(g):
    cc, dt, nstep, order, pusher, interp, friedman, Ns = g["meta_scalars"]
    kw = dict(Ns=int(Ns), cc=float(cc), delh=1.0, order=int(order), pusher=int(pusher), interp=int(interp),
              friedman=float(friedman), periodic=tuple(int(v) for v in g["meta_periodic"]))
    return tuple(int(v) for v in g["meta_ndims"]), tuple(int(v) for v in g["meta_cdims"]), kw, float(dt), int(nstep)


def upload(sim, g):
    for isp in range(sim.Ns):
        q, m = g[f"in_qm_{isp}"]
        sim.set_species(isp, float(q), float(m))
    for ic in range(sim.nchunk):
        sim.set_field(ic, 0, g[f"in_uf_{ic}"])
        for isp in range(sim.Ns):
            sim.set_particles(ic, isp, g[f"in_xu_{ic}_{isp}"])
    sim.finalize_setup()


def by_id(xu):
    return xu[np.argsort(xu[:, 6].view(np.int64), kind="stable")]


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)) if b.size else 0.0


def check_phases(make_sim, name, tol_u=1e-13, tol_x=1e-14, tol_j=1e-12, fused=False):
    """One step phase by phase against the reference's recorded outputs."""
    g = load(name)
    ndims, cdims, kw, dt, _ = config_of(g)
    sim = make_sim(ndims, cdims, **kw)
    upload(sim, g)
    sim.push_bfd(0.5 * dt)
    if fused:
        sim.push_deposit_fused(dt)
    else:
        sim.push_velocity(dt)
        for ic in range(sim.nchunk):
            for isp in range(sim.Ns):
                u = by_id(sim.get_particles(ic, isp))[:, 3:6]
                assert rel(u, g[f"p1_xu_{ic}_{isp}"]) < tol_u, (name, "velocity", ic, isp)
        sim.push_position(dt)
    for ic in range(sim.nchunk):
        for isp in range(sim.Ns):
            xu = sim.get_particles(ic, isp)
            order = np.argsort(xu[:, 6].view(np.int64), kind="stable")
            x = xu[order][:, 0:3]
            assert np.max(np.abs(x - g[f"p2_xu_{ic}_{isp}"])) < tol_x * max(ndims), (name, "position", ic, isp)
            # cell keys: bit-exact
            assert np.array_equal(sim.get_gindex(ic, isp)[order], g[f"p2_key_{ic}_{isp}"]), (name, "keys", ic, isp)
    if not fused:
        sim.deposit_current(dt)
    for ic in range(sim.nchunk):
        assert rel(sim.get_field(ic, 1), g[f"p3_f1_{ic}"]) < tol_j, (name, "current", ic)
    sim.exchange(1)
    for ic in range(sim.nchunk):
        assert rel(sim.get_field(ic, 1), g[f"p4_f1_{ic}"]) < tol_j, (name, "current halo", ic)
    sim.exchange(3)
    for ic in range(sim.nchunk):
        for isp in range(sim.Ns):
            assert np.array_equal(sim.get_pindex(ic, isp), g[f"p5_pindex_{ic}_{isp}"]), (name, "pindex", ic, isp)
    return sim


def check_multistep(make_sim, name, tol_f=1e-11, tol_p=1e-12):
    """nstep full steps: fields, current, phase space by id, pindex/Np bit-exact, residuals."""
    g = load(name)
    ndims, cdims, kw, dt, nstep = config_of(g)
    sim = make_sim(ndims, cdims, **kw)
    upload(sim, g)
    sim.step(dt, nstep)
    sim.synchronize()
    assert np.array_equal(sim.get_np_all(), g["end_np"]), (name, "Np")
    for ic in range(sim.nchunk):
        assert rel(sim.get_field(ic, 0), g[f"end_f0_{ic}"]) < tol_f, (name, "uf", ic)
        assert rel(sim.get_field(ic, 1), g[f"end_f1_{ic}"]) < tol_f, (name, "uj", ic)
        if f"end_f2_{ic}" in g.files:
            assert rel(sim.get_field(ic, 2)[..., :3], g[f"end_f2_{ic}"][..., :3]) < tol_f, (name, "ff", ic)
        for isp in range(sim.Ns):
            assert np.array_equal(sim.get_pindex(ic, isp), g[f"end_pindex_{ic}_{isp}"]), (name, "pindex", ic, isp)
            xu = by_id(sim.get_particles(ic, isp))
            ref = g[f"end_xu_{ic}_{isp}"]
            assert np.array_equal(xu[:, 6].view(np.int64), ref[:, 6].view(np.int64)), (name, "ids", ic, isp)
            assert np.max(np.abs(xu[:, 0:3] - ref[:, 0:3])) < tol_p * max(ndims), (name, "x", ic, isp)
            assert rel(xu[:, 3:6], ref[:, 3:6]) < tol_p, (name, "u", ic, isp)
    assert np.allclose(sim.get_diverror(), g["end_diverror"], atol=1e-10), (name, "diverror")
    return sim, g
