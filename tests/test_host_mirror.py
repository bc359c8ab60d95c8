"""The C++ host mirror (host/picnix_host.hpp): PicChunk-named per-chunk calls from OpenMP workers,
first caller launches the arena-wide kernel -- compiled with plain g++ against the C ABI only."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "host", "host_demo")


def build_demo():
    proc = subprocess.run(["make", "-C", os.path.join(ROOT, "host")], stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True)
    assert proc.returncode == 0, proc.stdout[-3000:]


def test_host_mirror_builds_and_fails_loudly_without_device():
    import torch

    build_demo()
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    proc = subprocess.run([DEMO, "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert proc.returncode == 3  # PICNIX_ERR_NODEVICE surfaced as picnix::host::Error, no CPU fallback
    assert "no CPU fallback" in proc.stderr or "picnix error 4" in proc.stderr


@pytest.mark.gpu
def test_host_mirror_push_openmp_equals_step():
    if not os.path.exists(DEMO):
        build_demo()
    env = dict(os.environ, OMP_NUM_THREADS="4")
    proc = subprocess.run([DEMO, "8"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    res = json.loads(proc.stdout.strip().splitlines()[-1])
    assert res["ok"] and res["same_np"] and res["kernel_launches"] > 0
    assert res["max_field_diff"] <= 1e-11 * res["field_scale"]
