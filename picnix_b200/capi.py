"""ctypes binding of libpicnix_b200.so -- the C ABI declared in include/picnix_b200.h.

The library is hand-written CUDA for sm_100a; there is no CPU fallback.  Importing this module
only loads the shared object (so symbol checks work on a box without a GPU); creating an arena
without a CUDA device fails loudly with PICNIX_ERR_NODEVICE.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpicnix_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_OVERFLOW, ERR_NODEVICE = 0, 1, 2, 3, 4
BC_NONE, BC_CONDUCTING, BC_WALL, BC_INFLOW = 0, 1, 2, 3
BOUNDARY_EMF, BOUNDARY_CUR, BOUNDARY_MOM, BOUNDARY_PARTICLE = 0, 1, 2, 3
FIELD_UF, FIELD_UJ, FIELD_FF, FIELD_UM = 0, 1, 2, 3
PUSHER_BORIS, PUSHER_VAY, PUSHER_HIGUERA_CARY = 0, 1, 2
INTERP_MC, INTERP_WT = 0, 1


class Config(C.Structure):
    """picnix_config_t"""

    _fields_ = [
        ("ndims", C.c_int32 * 3),
        ("cdims", C.c_int32 * 3),
        ("periodic", C.c_int32 * 3),
        ("order", C.c_int32),
        ("pusher", C.c_int32),
        ("interp", C.c_int32),
        ("Ns", C.c_int32),
        ("nrank", C.c_int32),
        ("rank", C.c_int32),
        ("cc", C.c_double),
        ("delx", C.c_double),
        ("dely", C.c_double),
        ("delz", C.c_double),
        ("friedman", C.c_double),
        ("buffer_ratio", C.c_double),
    ]


class PicnixError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"picnix_b200 status {status}: {message}")
        self.status = status


_vp, _i32, _i64, _dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
_pd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_pi = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

# name -> (restype, argtypes); every symbol include/picnix_b200.h declares
SIGNATURES = {
    "picnix_sfc_build": (_i32, [_i32, _i32, _i32, _pi, _pi]),
    "picnix_assign_initial": (_i32, [_pd, _i32, _i32, _pi]),
    "picnix_assign_rebalance": (_i32, [_pd, _i32, _i32, _pi]),
    "picnix_cuda_arena_create": (_i32, [C.POINTER(Config), C.c_void_p, C.POINTER(_vp)]),
    "picnix_cuda_arena_destroy": (_i32, [_vp]),
    "picnix_cuda_last_error": (C.c_char_p, [_vp]),
    "picnix_cuda_set_option": (_i32, [_vp, C.c_char_p, _i64]),
    "picnix_cuda_set_stream": (_i32, [_vp, _vp]),
    "picnix_cuda_synchronize": (_i32, [_vp]),
    "picnix_cuda_get_layout": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i32), _pi, C.POINTER(_i32),
                                      C.POINTER(_i32)]),
    "picnix_cuda_get_neighbors": (_i32, [_vp, _i32, _pi, _pi]),
    "picnix_cuda_set_species": (_i32, [_vp, _i32, _dbl, _dbl]),
    "picnix_cuda_set_particle_capacity": (_i32, [_vp, _pi]),
    "picnix_cuda_upload_field": (_i32, [_vp, _i32, _i32, _pd]),
    "picnix_cuda_download_field": (_i32, [_vp, _i32, _i32, _pd]),
    "picnix_cuda_upload_particles": (_i32, [_vp, _i32, _i32, _pd, _i32]),
    "picnix_cuda_download_particles": (_i32, [_vp, _i32, _i32, _i32, _i32, _pd]),
    "picnix_cuda_get_np": (_i32, [_vp, _pi]),
    "picnix_cuda_download_pindex": (_i32, [_vp, _i32, _i32, _pi]),
    "picnix_cuda_download_gindex": (_i32, [_vp, _i32, _i32, _i32, _pi]),
    "picnix_cuda_init_friedman": (_i32, [_vp, _i32, _i32]),
    "picnix_cuda_push_bfd": (_i32, [_vp, _i32, _i32, _dbl]),
    "picnix_cuda_push_efd": (_i32, [_vp, _i32, _i32, _dbl]),
    "picnix_cuda_push_velocity": (_i32, [_vp, _i32, _i32, _dbl]),
    "picnix_cuda_push_position": (_i32, [_vp, _i32, _i32, _dbl]),
    "picnix_cuda_deposit_current": (_i32, [_vp, _i32, _i32, _dbl]),
    "picnix_cuda_sort_particle": (_i32, [_vp, _i32, _i32]),
    "picnix_cuda_deposit_moment": (_i32, [_vp]),
    "picnix_cuda_get_particle_energy": (_i32, [_vp, _pd]),
    "picnix_cuda_push_deposit_fused": (_i32, [_vp, _i32, _i32, _dbl]),
    "picnix_cuda_boundary_begin": (_i32, [_vp, _i32]),
    "picnix_cuda_boundary_end": (_i32, [_vp, _i32]),
    "picnix_cuda_get_peers": (_i32, [_vp, C.POINTER(_i32), C.c_void_p]),
    "picnix_cuda_get_comm_buffer": (_i32, [_vp, _i32, _i32, C.POINTER(_vp), C.POINTER(_i64),
                                           C.POINTER(_vp), C.POINTER(_i64)]),
    "picnix_cuda_set_recv_bytes": (_i32, [_vp, _i32, _i32, _i64]),
    "picnix_cuda_step": (_i32, [_vp, _dbl, _i32]),
    "picnix_cuda_get_diverror": (_i32, [_vp, _pd, _pd]),
    "picnix_cuda_get_field_energy": (_i32, [_vp, _pd, _pd]),
    "picnix_cuda_get_counters": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "picnix_cuda_get_growth_stats": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "picnix_cuda_set_boundary_condition": (_i32, [_vp, _i32, _i32, _i32, C.c_void_p]),
    "picnix_cuda_inject_particles": (_i32, [_vp, _i32, _i32, _pd, _i32]),
    "picnix_cuda_step_host": (_i32, [_vp, _dbl, _i32, _pd, _pd, _pd, _pd, _pi, _pi, _pi]),
    "picnix_cuda_upload_state": (_i32, [_vp, _pd, _pd, _pd, _pd, _pi, _pi]),
    "picnix_cuda_download_state": (_i32, [_vp, _pd, _pd, _pd, _pd, _pi, _pi]),
    "picnix_cuda_chunk_pack_size": (_i32, [_vp, _i32, C.POINTER(_i64)]),
    "picnix_cuda_chunk_pack": (_i32, [_vp, _i32, _vp, _i64]),
    "picnix_cuda_chunk_unpack": (_i32, [_vp, _i32, _vp, _i64]),
    "picnix_cuda_host_alloc": (_i32, [C.POINTER(_vp), _i64]),
    "picnix_cuda_host_free": (_i32, [_vp]),
}

_lib = None


def load():
    """Load libpicnix_b200.so and attach the prototypes.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def sfc_build(Cz, Cy, Cx):
    """Chunk ordering of nix::ChunkMap: returns (chunkid[Cz,Cy,Cx], coord[n,3] as x,y,z)."""
    lib = load()
    chunkid = np.zeros((Cz, Cy, Cx), dtype=np.int32)
    coord = np.zeros((Cz * Cy * Cx, 3), dtype=np.int32)
    status = lib.picnix_sfc_build(Cz, Cy, Cx, chunkid.reshape(-1), coord.reshape(-1))
    if status != OK:
        raise PicnixError(status, "picnix_sfc_build")
    return chunkid, coord


def assign_initial(load_per_chunk, nrank):
    lib = load()
    loads = np.ascontiguousarray(load_per_chunk, dtype=np.float64)
    boundary = np.zeros(nrank + 1, dtype=np.int32)
    status = lib.picnix_assign_initial(loads, loads.size, nrank, boundary)
    if status != OK:
        raise PicnixError(status, "picnix_assign_initial")
    return boundary


def assign_rebalance(load_per_chunk, boundary):
    lib = load()
    loads = np.ascontiguousarray(load_per_chunk, dtype=np.float64)
    boundary = np.ascontiguousarray(boundary, dtype=np.int32).copy()
    status = lib.picnix_assign_rebalance(loads, loads.size, boundary.size - 1, boundary)
    if status != OK:
        raise PicnixError(status, "picnix_assign_rebalance")
    return boundary
