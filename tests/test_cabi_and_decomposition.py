"""CPU: the C-ABI library loads and exports every symbol include/picnix_b200.h declares; the host-side
integer logic (space-filling curve, neighbour tables, rank boundaries) is bit-exact with the oracle;
and the product fails loudly, with no CPU fallback, when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import port_backend
from picnix_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "picnix_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(picnix_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    names = declared_symbols()
    assert len(names) >= 38
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/picnix_b200.h but not exported"
        assert name in capi.SIGNATURES, f"{name} has no ctypes prototype in capi.SIGNATURES"
    for name in capi.SIGNATURES:
        assert name in names, f"{name} bound in capi.py but not declared in the header"


def test_product_does_not_reference_the_oracle():
    """No file of the product package may import, link or mention anything under oracle/."""
    pkg = os.path.join(ROOT, "picnix_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle." not in text.replace("oracle.\n", "") or "from oracle" not in text, fn
                assert "from oracle" not in text and "import oracle" not in text, fn
                assert "picnix_oracle" not in text and "ref_backend" not in text and "port_backend" not in text, fn


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from picnix_b200 import CudaSim

    with pytest.raises(capi.PicnixError) as err:
        CudaSim((8, 8, 8), (2, 2, 2), Ns=1, cc=1.0)
    assert err.value.status == capi.ERR_NODEVICE


@pytest.mark.parametrize("cdims", [(1, 1, 1), (1, 1, 7), (1, 1, 64), (1, 2, 2), (1, 4, 6), (1, 16, 16), (1, 6, 10),
                                   (1, 8, 2), (2, 2, 2), (4, 4, 4), (8, 8, 8), (4, 2, 6), (2, 6, 4), (6, 4, 2),
                                   (16, 16, 16), (2, 4, 16), (10, 6, 4), (3, 5, 7), (1, 3, 5)])
def test_sfc_bit_exact_and_valid(cdims):
    cid, coord = capi.sfc_build(*cdims)
    pid, pcoord = port_backend.sfc_build(*cdims)
    assert np.array_equal(cid, pid) and np.array_equal(coord, pcoord)
    n = cdims[0] * cdims[1] * cdims[2]
    # a permutation of 0..n-1 (nix/unittest/test_sfc.cpp:11-113)
    assert sorted(cid.reshape(-1).tolist()) == list(range(n))
    # coord is the inverse map, stored (x, y, z)
    for iz in range(cdims[0]):
        for iy in range(cdims[1]):
            for ix in range(cdims[2]):
                assert tuple(coord[cid[iz, iy, ix]]) == (ix, iy, iz)
    # locality: consecutive ids are neighbours; unit steps when all extents are even
    if n > 1:
        step = np.abs(np.diff(coord.astype(np.int64), axis=0))
        dist2 = (step ** 2).sum(axis=1)
        if all(c % 2 == 0 or c == 1 for c in cdims):
            assert dist2.max() == 1
        else:
            assert dist2.max() <= 3


def test_rank_boundaries_bit_exact():
    rng = np.random.default_rng(2)
    for nchunk, nrank in [(64, 8), (27, 4), (100, 7), (16, 16), (512, 8), (4096, 8), (9, 2)]:
        for trial in range(3):
            load = rng.uniform(0.1, 3.0, nchunk) if trial else np.ones(nchunk)
            b = capi.assign_initial(load, nrank)
            assert np.array_equal(b, port_backend.assign_initial(load, nrank))
            assert b[0] == 0 and b[-1] == nchunk and np.all(np.diff(b) > 0)
            load2 = load * rng.uniform(0.5, 1.5, nchunk)
            b2 = capi.assign_rebalance(load2, b)
            assert np.array_equal(b2, port_backend.assign_rebalance(load2, b))
            assert b2[0] == 0 and b2[-1] == nchunk and np.all(np.diff(b2) > 0)


def test_even_split_for_weak_scaling_layouts():
    """bench.py layouts: 512 chunks per GPU, contiguous SFC ranges, 8 ranks."""
    b = capi.assign_initial(np.ones(4096), 8)
    assert np.array_equal(b, np.arange(9) * 512)
