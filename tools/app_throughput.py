#!/usr/bin/env python
"""Throughput of the reference's OWN application (nix::Application::main -> PicApplication::push_openmp ->
example/thermal MainChunk, config.toml driven) with its chunks bound to the B200 library
(host/ref_binding/_build/thermal_cuda), at the benchmark size, and of the unmodified CPU application
(thermal_ref) on a bounded sample, both timed by the application's own log (log.msgpack: unix time stamp
of every step).

    python tools/app_throughput.py [--cells 128] [--steps 40] [--ref-cells 64] [--ref-steps 6]

PICNIX_SYNC_HOST_INTERVAL=0: the state stays resident, the host mirrors are refreshed only for
diagnostics (none in this run besides the history at the last step).  One JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import msgpack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "host", "ref_binding", "_build")

CONFIG = """
[application]
  basedir = 'data'
  [application.log]
    interval = 1
  [application.rebalance]
    interval = 1000000
  [application.option]
    vectorization = 'vector'
    seed_type = 'fixed'
    order = 2

[[diagnostic]]
  name = 'history'
  interval = {nstep}

[parameter]
  Nx = {n}
  Ny = {n}
  Nz = {n}
  Cx = {c}
  Cy = {c}
  Cz = {c}
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
    np = 32
    qm = -1.0
    ro = 1.0
    vt = 1.0

[[parameter.particle]]
    np = 32
    qm = +0.1
    ro = 10.0
    vt = 0.31622776601
"""


def run(binary, cells, nstep, env_extra):
    with tempfile.TemporaryDirectory() as work:
        with open(os.path.join(work, "config.toml"), "w") as fp:
            fp.write(CONFIG.format(n=cells, c=cells // 16, nstep=nstep))
        env = dict(os.environ, **env_extra)
        proc = subprocess.run([binary, "-c", "config.toml", "-t", str(0.05 * nstep)], cwd=work, env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout[-3000:])
            raise SystemExit(1)
        up = msgpack.Unpacker(raw=False)
        up.feed(open(os.path.join(work, "data", "log.msgpack"), "rb").read())
        recs = [r for r in up if isinstance(r, dict) and "step" in r]
        hist = open(os.path.join(work, "data", "history.txt")).read().strip().split("\n")[-1].split()
    recs.sort(key=lambda r: r["step"])
    skip = max(2, len(recs) // 5)                      # warm-up steps
    t0, t1 = recs[skip]["timestamp"]["unixtime"], recs[-1]["timestamp"]["unixtime"]
    nst = recs[-1]["step"] - recs[skip]["step"]
    return (t1 - t0) / nst, nst, [float(v) for v in hist]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=128)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--ref-cells", type=int, default=64)
    ap.add_argument("--ref-steps", type=int, default=6)
    args = ap.parse_args()
    ncore = len(os.sched_getaffinity(0))
    sec, nst, hist = run(os.path.join(BUILD, "thermal_cuda"), args.cells, args.steps,
                         dict(OMP_NUM_THREADS="8", PICNIX_SYNC_HOST_INTERVAL="0"))
    npart = args.cells ** 3 * 64
    out = {"application": "nix::Application::main + example/thermal/main.cpp (unmodified), chunks = CudaPicChunk",
           "cells": args.cells ** 3, "chunks": (args.cells // 16) ** 3, "particles": npart, "timed_steps": nst,
           "ms_per_step": 1e3 * sec, "particle_steps_per_s": npart / sec,
           "last_history_row": hist}
    rsec, rnst, _ = run(os.path.join(BUILD, "thermal_ref"), args.ref_cells, args.ref_steps,
                        dict(OMP_NUM_THREADS=str(ncore)))
    rpart = args.ref_cells ** 3 * 64
    out["cpu_application"] = {"cells": args.ref_cells ** 3, "particles": rpart, "timed_steps": rnst, "cores": ncore,
                              "ms_per_step": 1e3 * rsec, "particle_steps_per_s": rpart / rsec}
    out["ratio"] = out["particle_steps_per_s"] / out["cpu_application"]["particle_steps_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
