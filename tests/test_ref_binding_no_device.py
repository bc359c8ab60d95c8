"""CPU: the reference's application bound to the CUDA library (host/ref_binding/_build/*_cuda) must fail
loudly on a machine without a GPU -- there is no CPU fallback behind CudaPicChunk -- while the unmodified
application built next to it from the same sources runs."""
import os
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "host", "ref_binding", "_build")

CONFIG = """
[application]
  basedir = 'data'
  [application.rebalance]
    interval = 1000000
  [application.option]
    seed_type = 'fixed'
[[diagnostic]]
  name = 'history'
  interval = 1
[parameter]
  Nx = 16
  Ny = 16
  Nz = 1
  Cx = 2
  Cy = 2
  Cz = 1
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
    np = 4
    qm = -1.0
    ro = 1.0
    vt = 1.0
[[parameter.particle]]
    np = 4
    qm = +0.1
    ro = 10.0
    vt = 0.31622776601
"""


def run(binary, workdir):
    os.makedirs(workdir, exist_ok=True)
    with open(os.path.join(workdir, "config.toml"), "w") as fp:
        fp.write(CONFIG)
    return subprocess.run([binary, "-c", "config.toml", "-t", "0.2"], cwd=workdir, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="2"))


@pytest.mark.skipif(not os.path.exists(os.path.join(BUILD, "thermal_cuda")), reason="host/ref_binding/_build not built")
@pytest.mark.skipif(torch.cuda.is_available(), reason="this test is about machines without a GPU")
def test_bound_application_refuses_to_run_without_a_device(tmp_path):
    ref = run(os.path.join(BUILD, "thermal_ref"), str(tmp_path / "ref"))
    assert ref.returncode == 0, ref.stdout[-2000:]
    assert os.path.exists(tmp_path / "ref" / "data" / "history.txt")
    gpu = run(os.path.join(BUILD, "thermal_cuda"), str(tmp_path / "gpu"))
    assert gpu.returncode != 0
    assert "no CPU fallback" in gpu.stdout
