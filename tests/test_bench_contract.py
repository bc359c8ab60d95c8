"""CPU: the host-side logic of bench.py -- workload table, per-N box layout, and the reference arm's
JSON line (the driver parses exactly one line; keys per the measurement contract)."""
import json
import os
import subprocess
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import ref_backend  # noqa: E402


def args(**kw):
    base = dict(cells=0, ppc=0, ref_cells=0, parity_cells=0)
    base.update(kw)
    return types.SimpleNamespace(**base)


@pytest.mark.parametrize("name", ["thermal3d", "twostream", "cherenkov"])
def test_workload_boxes_and_layouts(name):
    wl = bench.Workload(name, args())
    assert wl.Ns == len(wl.species) == len(wl.ppc)
    for ngpu in (1, 2, 4, 8):
        lay = wl.layout(ngpu)
        assert int(np.prod(lay)) == ngpu
        for per_rank in (wl.dims, wl.parity):
            ndims, cdims = wl.box(per_rank, ngpu)
            for n, c, k, l, d in zip(ndims, cdims, wl.chunk, lay, per_rank):
                assert n == d * l and c * k == n          # whole chunks, blocks of per-rank extent
                assert l == 1 or d > 1                   # ranks are only laid out along real dimensions
    ndims, cdims = wl.box(wl.sample)
    assert all(c * k == n for n, c, k in zip(ndims, cdims, wl.chunk))
    cfg = bench.workload_config(wl, 8)
    assert cfg["particles_per_gpu"] == int(np.prod(wl.dims)) * sum(wl.ppc)
    assert set(cfg) >= {"workload", "cells_per_gpu", "chunks_per_gpu", "parallelism", "l2_policy"}


def test_headline_workload_is_the_survey_configuration():
    wl = bench.Workload("thermal3d", args())
    assert wl.dims == (128, 128, 128) and wl.chunk == (16, 16, 16) and wl.ppc == (32, 32)
    assert (wl.cc, wl.delt, wl.delh, wl.B0) == (10.0, 0.05, 1.0, (5.0, 0.0, 0.0))
    assert wl.metric == bench.METRIC


def test_ncu_capture_lookup_matches_workload_sizes():
    for name in ("thermal3d", "cherenkov", "twostream"):
        wl = bench.Workload(name, args())
        ncell = int(np.prod(wl.dims))
        traffic, pipes = bench.ncu_traffic(ncell * sum(wl.ppc), ncell)
        alg = ncell * sum(wl.ppc) * bench.BYTES_PUSH_PER_PARTICLE + ncell * bench.BYTES_PUSH_PER_CELL
        assert traffic is not None and 1.0 <= traffic / alg < 1.1      # no wasted DRAM traffic
        assert pipes["warps_per_sm"] in (8, 12, 16)
    assert bench.ncu_traffic(12345, 678) == (None, None)


@pytest.mark.skipif(not ref_backend.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("workload,cells", [("thermal3d", 16), ("twostream", 256), ("cherenkov", 32)])
def test_reference_arm_prints_one_contract_line(workload, cells):
    env = dict(os.environ, OMP_NUM_THREADS="1")          # what torchrun exports: must be overridden
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                          "--steps", "2", "--warmup", "1", "--ref-cells", str(cells)],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1
    assert d["unit"] == bench.UNIT and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "THIS ARM: CPU sample" in d["config"]["workload"] and d["config"]["sample_particles"] > 0
    assert d["metric"] == bench.Workload(workload, args()).metric
