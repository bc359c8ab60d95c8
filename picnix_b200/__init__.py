"""picnix_b200 -- B200-native (sm_100a) implementation of the PIC-NIX per-timestep hot path.

Layout
------
csrc/            hand-written CUDA kernels + the C ABI (include/picnix_b200.h) -> libpicnix_b200.so
host/            C++17 host-side mirror of the reference's PicChunk / PicApplication interface
capi.py          ctypes binding of the C ABI
simulation.py    Python mirror of the chunk-level API (used by tests and bench.py)
problems.py      synthetic workloads of BASELINE.json (thermal plasma, beams)
distributed.py   one-process-per-GPU driver (torch.distributed transport between arenas)
"""
from . import capi  # noqa: F401
from .simulation import CudaSim  # noqa: F401
