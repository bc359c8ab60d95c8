"""ctypes wrapper around oracle/libpicnix_oracle.so, the plain-C restatement (TEST INFRASTRUCTURE ONLY).

`PortSim` has the same chunk-level API as `ref_backend.RefSim` (the compiled reference) and
`picnix_b200.CudaSim` (the product), plus the multi-rank plumbing (`peers`, `comm_buffer`,
`set_recv_bytes`) with HOST buffers so that `picnix_b200.distributed.Transport` can be exercised
with the gloo backend on a CPU-only box.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference leg may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpicnix_oracle.so")

MODE_EMF, MODE_CUR, MODE_MOM, MODE_PARTICLE = 0, 1, 2, 3
FIELD_UF, FIELD_UJ, FIELD_FF, FIELD_UM = 0, 1, 2, 3


class OrcConfig(C.Structure):
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
        ("simd_width", C.c_int32),
        ("nthread", C.c_int32),
        ("cc", C.c_double),
        ("delx", C.c_double),
        ("dely", C.c_double),
        ("delz", C.c_double),
        ("friedman", C.c_double),
        ("buffer_ratio", C.c_double),
    ]


_lib = None


def available():
    return os.path.exists(LIB_PATH) or os.path.exists(os.path.join(_HERE, "picnix_oracle.c"))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # plain C, gcc only: build on demand (this is the checker, not the product)
        subprocess.run(["make", "-C", _HERE, "port"], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    pd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    pi = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    sig = {
        "orc_sfc_build": (i32, [i32, i32, i32, pi, pi]),
        "orc_assign_initial": (i32, [pd, i32, i32, pi]),
        "orc_assign_rebalance": (i32, [pd, i32, i32, pi]),
        "orc_create": (vp, [C.POINTER(OrcConfig), vp]),
        "orc_destroy": (None, [vp]),
        "orc_num_chunks": (i32, [vp]),
        "orc_num_threads": (i32, [vp]),
        "orc_chunk_id_begin": (i32, [vp]),
        "orc_get_shape": (None, [vp, pi]),
        "orc_get_neighbors": (None, [vp, i32, pi, pi]),
        "orc_set_species": (None, [vp, i32, dbl, dbl]),
        "orc_set_field": (None, [vp, i32, i32, pd]),
        "orc_get_field": (None, [vp, i32, i32, pd]),
        "orc_set_particles": (None, [vp, i32, i32, pd, i32, i32]),
        "orc_get_np": (i32, [vp, i32, i32]),
        "orc_get_particles": (None, [vp, i32, i32, i32, i32, pd]),
        "orc_get_pindex": (None, [vp, i32, i32, pi]),
        "orc_get_gindex": (None, [vp, i32, i32, i32, pi]),
        "orc_init_friedman": (None, [vp]),
        "orc_push_bfd": (None, [vp, dbl]),
        "orc_push_efd": (None, [vp, dbl]),
        "orc_push_velocity": (None, [vp, dbl]),
        "orc_push_position": (None, [vp, dbl]),
        "orc_deposit_current": (None, [vp, dbl]),
        "orc_deposit_moment": (None, [vp]),
        "orc_sort_particle": (None, [vp]),
        "orc_boundary_begin": (None, [vp, i32]),
        "orc_boundary_end": (None, [vp, i32]),
        "orc_get_peers": (i32, [vp, vp]),
        "orc_get_comm_buffer": (None, [vp, i32, i32, C.POINTER(vp), C.POINTER(i64), C.POINTER(vp), C.POINTER(i64)]),
        "orc_set_recv_bytes": (None, [vp, i32, i32, i64]),
        "orc_exchange": (None, [vp, i32]),
        "orc_step": (None, [vp, dbl, i32]),
        "orc_get_diverror": (None, [vp, i32, C.POINTER(dbl), C.POINTER(dbl)]),
        "orc_get_energy": (None, [vp, i32, C.POINTER(dbl), C.POINTER(dbl), pd]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def sfc_build(Cz, Cy, Cx):
    lib = load()
    chunkid = np.zeros((Cz, Cy, Cx), dtype=np.int32)
    coord = np.zeros((Cz * Cy * Cx, 3), dtype=np.int32)
    lib.orc_sfc_build(Cz, Cy, Cx, chunkid.reshape(-1), coord.reshape(-1))
    return chunkid, coord


def assign_initial(load_per_chunk, nrank):
    lib = load()
    loads = np.ascontiguousarray(load_per_chunk, dtype=np.float64)
    boundary = np.zeros(nrank + 1, dtype=np.int32)
    lib.orc_assign_initial(loads, loads.size, nrank, boundary)
    return boundary


def assign_rebalance(load_per_chunk, boundary):
    lib = load()
    loads = np.ascontiguousarray(load_per_chunk, dtype=np.float64)
    boundary = np.ascontiguousarray(boundary, dtype=np.int32).copy()
    lib.orc_assign_rebalance(loads, loads.size, boundary.size - 1, boundary)
    return boundary


class PortSim:
    """All chunks of one rank held by the C restatement."""

    name = "port"

    def __init__(self, ndims, cdims, Ns, cc, delh=1.0, order=2, pusher=0, interp=0, periodic=(1, 1, 1),
                 friedman=0.0, buffer_ratio=0.2, nrank=1, rank=0, boundary=None, simd_width=8, nthread=0):
        self.lib = load()
        cfg = OrcConfig()
        cfg.ndims[:] = ndims
        cfg.cdims[:] = cdims
        cfg.periodic[:] = periodic
        cfg.order, cfg.pusher, cfg.interp, cfg.Ns = order, pusher, interp, Ns
        cfg.nrank, cfg.rank, cfg.simd_width, cfg.nthread = nrank, rank, simd_width, nthread
        cfg.cc = cc
        if np.isscalar(delh):
            cfg.delx = cfg.dely = cfg.delz = delh
        else:
            cfg.delz, cfg.dely, cfg.delx = delh
        cfg.friedman, cfg.buffer_ratio = friedman, buffer_ratio
        self.cfg = cfg
        bptr = None
        if boundary is not None:
            self._boundary = np.ascontiguousarray(boundary, dtype=np.int32)
            bptr = self._boundary.ctypes.data_as(C.c_void_p)
        self.h = self.lib.orc_create(C.byref(cfg), bptr)
        if not self.h:
            raise ValueError("orc_create: invalid configuration")
        self.Ns = Ns
        self.nchunk = self.lib.orc_num_chunks(self.h)
        self.chunk_id_begin = self.lib.orc_chunk_id_begin(self.h)
        shape = np.zeros(5, dtype=np.int32)
        self.lib.orc_get_shape(self.h, shape)
        self.shape = tuple(int(v) for v in shape[:3])
        self.nb, self.Ng = int(shape[3]), int(shape[4])
        self.nthread = self.lib.orc_num_threads(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        pass

    def commit(self):
        pass

    # -- decomposition ---------------------------------------------------------------------
    def chunkmap(self):
        return sfc_build(*tuple(self.cfg.cdims))

    def neighbors(self, ic):
        nbid = np.zeros(27, dtype=np.int32)
        nbrank = np.zeros(27, dtype=np.int32)
        self.lib.orc_get_neighbors(self.h, ic, nbid, nbrank)
        return nbid, nbrank

    # -- state -----------------------------------------------------------------------------
    def _field_shape(self, which):
        tail = {FIELD_UF: (6,), FIELD_UJ: (4,), FIELD_FF: (3, 6), FIELD_UM: (self.Ns, 14)}[which]
        return self.shape + tail

    def set_species(self, isp, q, m):
        self.lib.orc_set_species(self.h, isp, q, m)

    def set_field(self, ic, which, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        assert arr.shape == self._field_shape(which), (arr.shape, self._field_shape(which))
        self.lib.orc_set_field(self.h, ic, which, arr.reshape(-1))

    def get_field(self, ic, which):
        out = np.zeros(self._field_shape(which), dtype=np.float64)
        self.lib.orc_get_field(self.h, ic, which, out.reshape(-1))
        return out

    def set_particles(self, ic, isp, xu, np_alloc=None):
        xu = np.ascontiguousarray(xu, dtype=np.float64).reshape(-1, 7)
        n = xu.shape[0]
        if np_alloc is None:
            np_alloc = int(n * (1 + self.cfg.buffer_ratio))
        self.lib.orc_set_particles(self.h, ic, isp, xu.reshape(-1), n, np_alloc)

    def get_np(self, ic, isp):
        return self.lib.orc_get_np(self.h, ic, isp)

    def get_np_all(self):
        return np.array([[self.get_np(ic, isp) for isp in range(self.Ns)] for ic in range(self.nchunk)],
                        dtype=np.int32)

    def get_particles(self, ic, isp, which=0, n=None):
        if n is None:
            n = self.get_np(ic, isp)
        out = np.zeros((n, 7), dtype=np.float64)
        if n > 0:
            self.lib.orc_get_particles(self.h, ic, isp, which, n, out.reshape(-1))
        return out

    def get_pindex(self, ic, isp):
        out = np.zeros(self.Ng + 1, dtype=np.int32)
        self.lib.orc_get_pindex(self.h, ic, isp, out)
        return out

    def get_gindex(self, ic, isp, n=None):
        if n is None:
            n = self.get_np(ic, isp)
        out = np.zeros(max(n, 1), dtype=np.int32)
        if n > 0:
            self.lib.orc_get_gindex(self.h, ic, isp, n, out)
        return out[:n]

    # -- phases ----------------------------------------------------------------------------
    def finalize_setup(self):
        """Tail of MainChunk::setup + PicApplication::setup_chunks (pic/pic_application.cpp:106-130)."""
        self.init_friedman()
        self.sort_particle()
        self.exchange(MODE_EMF)

    def init_friedman(self):
        self.lib.orc_init_friedman(self.h)

    def push_bfd(self, dt):
        self.lib.orc_push_bfd(self.h, dt)

    def push_efd(self, dt):
        self.lib.orc_push_efd(self.h, dt)

    def push_velocity(self, dt):
        self.lib.orc_push_velocity(self.h, dt)

    def push_position(self, dt):
        self.lib.orc_push_position(self.h, dt)

    def deposit_current(self, dt):
        self.lib.orc_deposit_current(self.h, dt)

    def push_deposit_fused(self, dt):
        self.push_velocity(dt)
        self.push_position(dt)
        self.deposit_current(dt)

    def deposit_moment(self):
        self.lib.orc_deposit_moment(self.h)

    def sort_particle(self):
        self.lib.orc_sort_particle(self.h)

    def boundary_begin(self, mode):
        self.lib.orc_boundary_begin(self.h, mode)

    def boundary_end(self, mode):
        self.lib.orc_boundary_end(self.h, mode)

    def exchange(self, mode):
        if self.cfg.nrank != 1:
            raise RuntimeError("exchange() on a multi-rank PortSim needs a transport")
        self.lib.orc_exchange(self.h, mode)

    def step(self, dt, nstep=1):
        if self.cfg.nrank != 1:
            raise RuntimeError("step() on a multi-rank PortSim needs a transport")
        self.lib.orc_step(self.h, dt, nstep)

    def get_diverror(self):
        e, b = C.c_double(), C.c_double()
        out = np.zeros((self.nchunk, 2))
        for ic in range(self.nchunk):
            self.lib.orc_get_diverror(self.h, ic, C.byref(e), C.byref(b))
            out[ic] = e.value, b.value
        return out

    def get_energy(self):
        e, b = C.c_double(), C.c_double()
        out = np.zeros((self.nchunk, 2 + self.Ns))
        p = np.zeros(self.Ns)
        for ic in range(self.nchunk):
            self.lib.orc_get_energy(self.h, ic, C.byref(e), C.byref(b), p)
            out[ic, 0], out[ic, 1] = e.value, b.value
            out[ic, 2:] = p
        return out

    def get_field_energy(self):
        return self.get_energy()[:, :2]

    # -- multi-rank plumbing (host buffers) ------------------------------------------------
    def peers(self):
        n = self.lib.orc_get_peers(self.h, None)
        ranks = np.zeros(max(n, 1), dtype=np.int32)
        self.lib.orc_get_peers(self.h, ranks.ctypes.data_as(C.c_void_p))
        return [int(r) for r in ranks[:n]]

    def comm_buffer(self, mode, peer_index):
        sp, rp = C.c_void_p(), C.c_void_p()
        sb, rb = C.c_int64(), C.c_int64()
        self.lib.orc_get_comm_buffer(self.h, mode, peer_index, C.byref(sp), C.byref(sb), C.byref(rp), C.byref(rb))
        return sp.value, sb.value, rp.value, rb.value

    def set_recv_bytes(self, mode, peer_index, nbytes):
        self.lib.orc_set_recv_bytes(self.h, mode, peer_index, nbytes)
