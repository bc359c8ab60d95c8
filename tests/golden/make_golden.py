#!/usr/bin/env python
"""Generate the golden vectors in tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref).

Run in the container where /root/reference exists, after `make -C oracle ref`:

    python tests/golden/make_golden.py

Every file holds the inputs (fields and particles of every chunk, exactly as uploaded) and the
reference's outputs: per-phase results of ONE step (velocity, position + keys, current before and
after the halo add) and the state after `nstep` full steps (fields, current, particles sorted by
id, pindex, Np, divergence residuals, energies).  The fixtures travel with the repository; the
reference does not.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_backend  # noqa: E402
from picnix_b200 import problems  # noqa: E402

CASES = {
    # name: ndims, cdims, species, ppc, cc, B0, dt, nstep, order, pusher, interp, friedman, periodic
    "thermal3d_o2": ((8, 8, 8), (2, 2, 2), "thermal", (2, 2), 10.0, (5.0, 0.0, 0.0), 0.05, 8, 2, 0, 0, 0.0, (1, 1, 1)),
    "thermal3d_o1_wt_vay": ((8, 8, 8), (2, 2, 2), "thermal", (2, 2), 10.0, (5.0, 1.0, 0.5), 0.05, 4, 1, 1, 1, 0.1,
                            (1, 1, 1)),
    "thermal3d_o3_open": ((8, 8, 8), (2, 2, 2), "thermal", (2, 2), 10.0, (5.0, 0.0, 0.0), 0.05, 4, 3, 2, 0, 0.0,
                          (1, 1, 0)),
    "thermal2d_o2": ((1, 16, 16), (1, 2, 2), "thermal", (4, 4), 10.0, (5.0, 0.0, 0.0), 0.05, 8, 2, 0, 0, 0.05,
                     (1, 1, 1)),
    "twostream1d_o2": ((1, 1, 64), (1, 1, 8), "twostream", (8, 8, 16), 50.0, (10.0, 0.0, 0.0), 0.01, 20, 2, 0, 0,
                       0.0, (1, 1, 1)),
    "thermal1d_o4": ((1, 1, 32), (1, 1, 4), "thermal", (8, 8), 10.0, (2.0, 1.0, 0.0), 0.05, 6, 4, 0, 1, 0.0,
                     (1, 1, 1)),
}
SPECIES = {"thermal": problems.THERMAL_SPECIES, "twostream": problems.TWOSTREAM_SPECIES}


def sorted_by_id(xu):
    return xu[np.argsort(xu[:, 6].view(np.int64), kind="stable")]


class Recorder:
    """Backend shim that records what setup_uniform_plasma uploads."""

    def __init__(self, sim):
        self.sim, self.inputs = sim, {}
        self.nchunk, self.shape, self.nb, self.Ns = sim.nchunk, sim.shape, sim.nb, sim.Ns

    def chunkmap(self):
        return self.sim.chunkmap()

    def set_species(self, isp, q, m):
        self.inputs[f"qm_{isp}"] = np.array([q, m])
        self.sim.set_species(isp, q, m)

    def set_field(self, ic, which, arr):
        self.inputs[f"uf_{ic}"] = np.array(arr)
        self.sim.set_field(ic, which, arr)

    def set_particles(self, ic, isp, xu):
        self.inputs[f"xu_{ic}_{isp}"] = np.array(xu)
        self.sim.set_particles(ic, isp, xu)

    def finalize_setup(self):
        self.sim.finalize_setup()


def snapshot(sim, prefix, out, fields=(0, 1), particles=True, which=0, cols=slice(0, 7)):
    for ic in range(sim.nchunk):
        for f in fields:
            out[f"{prefix}_f{f}_{ic}"] = sim.get_field(ic, f)
        if particles:
            for isp in range(sim.Ns):
                out[f"{prefix}_xu_{ic}_{isp}"] = sorted_by_id(sim.get_particles(ic, isp, which))[:, cols].copy()


def main():
    lib = ref_backend.library_path()
    print("reference library:", lib)
    for name, (ndims, cdims, spname, ppc, cc, B0, dt, nstep, order, pusher, interp, friedman, periodic) in \
            CASES.items():
        species = SPECIES[spname]
        kw = dict(Ns=len(species), cc=cc, delh=1.0, order=order, pusher=pusher, interp=interp, friedman=friedman,
                  periodic=periodic)

        def fresh():
            sim = ref_backend.RefSim(ndims, cdims, vector_mode=1, **kw)
            rec = Recorder(sim)
            problems.setup_uniform_plasma(rec, ndims, cdims, species, ppc, B0=B0, seed=11, perturb=0.01)
            return sim, rec

        out = {}
        sim, rec = fresh()
        out.update({"in_" + k: v for k, v in rec.inputs.items()})
        out["meta_ndims"], out["meta_cdims"] = np.array(ndims), np.array(cdims)
        out["meta_ppc"], out["meta_periodic"] = np.array(ppc), np.array(periodic)
        out["meta_scalars"] = np.array([cc, dt, nstep, order, pusher, interp, friedman, len(species)], dtype=np.float64)
        out["meta_species"] = np.array(spname)
        # ---- one step, phase by phase ----
        sim.push_bfd(0.5 * dt)
        sim.push_velocity(dt)
        snapshot(sim, "p1", out, fields=(), cols=slice(3, 6))
        sim.push_position(dt)
        snapshot(sim, "p2", out, fields=(), cols=slice(0, 3))
        for ic in range(sim.nchunk):
            for isp in range(sim.Ns):
                xu = sim.get_particles(ic, isp)
                order_ = np.argsort(xu[:, 6].view(np.int64), kind="stable")
                out[f"p2_key_{ic}_{isp}"] = sim.get_gindex(ic, isp)[order_]
        sim.deposit_current(dt)
        snapshot(sim, "p3", out, fields=(1,), particles=False)
        sim.exchange(1)
        snapshot(sim, "p4", out, fields=(1,), particles=False)
        sim.exchange(3)
        for ic in range(sim.nchunk):
            for isp in range(sim.Ns):
                out[f"p5_pindex_{ic}_{isp}"] = sim.get_pindex(ic, isp)
        # ---- nstep full steps from the same inputs ----
        sim2, _ = fresh()
        sim2.step(dt, nstep)
        snapshot(sim2, "end", out, fields=(0, 1, 2) if friedman != 0.0 else (0, 1))
        for ic in range(sim2.nchunk):
            for isp in range(sim2.Ns):
                out[f"end_pindex_{ic}_{isp}"] = sim2.get_pindex(ic, isp)
        out["end_np"] = np.array([[sim2.get_np(ic, isp) for isp in range(sim2.Ns)] for ic in range(sim2.nchunk)])
        out["end_diverror"] = sim2.get_diverror()
        sim2.deposit_moment()
        out["end_energy"] = sim2.get_energy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, np = {int(out['end_np'].sum())}")


if __name__ == "__main__":
    main()
