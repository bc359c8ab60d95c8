"""ctypes wrapper around oracle/_ref/libpicnix_ref_*.so (TEST INFRASTRUCTURE ONLY).

The shared library is the UNMODIFIED reference (amanotk/pic-nix: pic/pic_chunk.cpp, nix/chunk.cpp,
nix/chunkmap.cpp, nix/sfc.cpp) compiled by oracle/Makefile against the single-process MPI shim;
oracle/ref_driver.cpp exposes its PicChunk entry points (pic/pic_chunk.hpp:90-143) as a flat C API.

Only tests/, __graft_entry__.smoke() and bench.py's reference arm may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

MODE_EMF, MODE_CUR, MODE_MOM, MODE_PARTICLE = 0, 1, 2, 3
FIELD_UF, FIELD_UJ, FIELD_FF, FIELD_UM = 0, 1, 2, 3


class RefConfig(C.Structure):
    _fields_ = [
        ("ndims", C.c_int32 * 3),
        ("cdims", C.c_int32 * 3),
        ("periodic", C.c_int32 * 3),
        ("order", C.c_int32),
        ("pusher", C.c_int32),
        ("interp", C.c_int32),
        ("Ns", C.c_int32),
        ("vector_mode", C.c_int32),
        ("nthread", C.c_int32),
        ("problem", C.c_int32),
        ("cc", C.c_double),
        ("delh", C.c_double),
        ("friedman", C.c_double),
        ("buffer_ratio", C.c_double),
    ]


def _has_avx512():
    try:
        with open("/proc/cpuinfo") as fp:
            for line in fp:
                if line.startswith("flags"):
                    return " avx512f " in line + " "
    except OSError:
        pass
    return False


def library_path(prefer_wide=True):
    """Return the path of the best prebuilt reference library for this host (or None)."""
    names = []
    if prefer_wide and _has_avx512():
        names.append("libpicnix_ref_x86-64-v4.so")
    names.append("libpicnix_ref_x86-64-v3.so")
    for name in names:
        path = os.path.join(_HERE, "_ref", name)
        if os.path.exists(path):
            return path
    return None


def available():
    return library_path() is not None


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if path is None:
        raise RuntimeError("oracle/_ref is not built (run `make -C oracle ref` where /root/reference exists)")
    lib = C.CDLL(path)
    vp, i32, dbl = C.c_void_p, C.c_int32, C.c_double
    pd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    pi = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    sig = {
        "ref_create": (vp, [C.POINTER(RefConfig)]),
        "ref_destroy": (None, [vp]),
        "ref_num_chunks": (i32, [vp]),
        "ref_num_threads": (i32, [vp]),
        "ref_get_shape": (None, [vp, pi]),
        "ref_get_chunkmap": (None, [vp, pi, pi]),
        "ref_get_neighbors": (None, [vp, i32, pi, pi]),
        "ref_chunkmap_validate": (i32, [vp]),
        "ref_set_species": (None, [vp, i32, dbl, dbl]),
        "ref_get_species": (None, [vp, i32, C.POINTER(dbl), C.POINTER(dbl)]),
        "ref_set_problem_json": (None, [C.c_char_p]),
        "ref_set_field": (None, [vp, i32, i32, pd]),
        "ref_get_field": (None, [vp, i32, i32, pd]),
        "ref_set_particles": (None, [vp, i32, i32, pd, i32, i32]),
        "ref_get_np": (i32, [vp, i32, i32]),
        "ref_get_np_total": (i32, [vp, i32, i32]),
        "ref_get_particles": (None, [vp, i32, i32, i32, i32, pd]),
        "ref_get_pindex": (None, [vp, i32, i32, pi]),
        "ref_get_gindex": (None, [vp, i32, i32, i32, pi]),
        "ref_finalize_setup": (None, [vp]),
        "ref_init_friedman": (None, [vp]),
        "ref_push_bfd": (None, [vp, dbl]),
        "ref_push_efd": (None, [vp, dbl]),
        "ref_push_velocity": (None, [vp, dbl]),
        "ref_push_position": (None, [vp, dbl]),
        "ref_deposit_current": (None, [vp, dbl]),
        "ref_deposit_moment": (None, [vp]),
        "ref_sort_particle": (None, [vp]),
        "ref_exchange": (None, [vp, i32]),
        "ref_get_diverror": (None, [vp, i32, C.POINTER(dbl), C.POINTER(dbl)]),
        "ref_get_energy": (None, [vp, i32, C.POINTER(dbl), C.POINTER(dbl), pd]),
        "ref_step": (None, [vp, dbl, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class RefSim:
    """All chunks of a (small) run held by the reference's own PicChunk objects in this process."""

    name = "reference"

    PROBLEMS = {None: 0, "plain": 0, "mrx": 1, "shock": 2}

    def __init__(self, ndims, cdims, Ns, cc, delh=1.0, order=2, pusher=0, interp=0, periodic=(1, 1, 1),
                 friedman=0.0, buffer_ratio=0.2, vector_mode=1, nthread=0, problem=None, problem_json=None):
        """problem: None / "mrx" / "shock" -- the chunks are the example's MainChunk, i.e. its physical
        boundary hooks (set_boundary_field / set_boundary_particle / inject_particle) are active.
        problem_json: dict merged into the chunk configuration; with "example_setup": True the example's
        own MainChunk::setup builds the initial state (Harris sheet, shock tube), otherwise the state
        comes through set_field / set_particles as usual."""
        self.lib = load()
        if problem_json is not None:
            import json

            self.lib.ref_set_problem_json(json.dumps(problem_json).encode())
        cfg = RefConfig()
        cfg.problem = self.PROBLEMS[problem]
        cfg.ndims[:] = ndims
        cfg.cdims[:] = cdims
        cfg.periodic[:] = periodic
        cfg.order, cfg.pusher, cfg.interp, cfg.Ns = order, pusher, interp, Ns
        cfg.vector_mode, cfg.nthread = vector_mode, nthread
        cfg.cc, cfg.delh, cfg.friedman, cfg.buffer_ratio = cc, delh, friedman, buffer_ratio
        self.cfg = cfg
        self.h = self.lib.ref_create(C.byref(cfg))
        self.Ns = Ns
        self.nchunk = self.lib.ref_num_chunks(self.h)
        shape = np.zeros(5, dtype=np.int32)
        self.lib.ref_get_shape(self.h, shape)
        self.shape = tuple(int(s) for s in shape[:3])
        self.nb = int(shape[3])
        self.Ng = int(shape[4])
        self.nthread = self.lib.ref_num_threads(self.h)

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- decomposition ---------------------------------------------------------------------
    def chunkmap(self):
        cd = tuple(self.cfg.cdims)
        chunkid = np.zeros(cd, dtype=np.int32)
        coord = np.zeros((self.nchunk, 3), dtype=np.int32)
        self.lib.ref_get_chunkmap(self.h, chunkid.reshape(-1), coord.reshape(-1))
        return chunkid, coord

    def neighbors(self, ic):
        nbid = np.zeros(27, dtype=np.int32)
        nbrank = np.zeros(27, dtype=np.int32)
        self.lib.ref_get_neighbors(self.h, ic, nbid, nbrank)
        return nbid, nbrank

    def chunkmap_validate(self):
        return bool(self.lib.ref_chunkmap_validate(self.h))

    # -- state -----------------------------------------------------------------------------
    def _field_shape(self, which):
        tail = {FIELD_UF: (6,), FIELD_UJ: (4,), FIELD_FF: (3, 6), FIELD_UM: (self.Ns, 14)}[which]
        return self.shape + tail

    def set_species(self, isp, q, m):
        self.lib.ref_set_species(self.h, isp, q, m)

    def set_field(self, ic, which, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        assert arr.shape == self._field_shape(which), (arr.shape, self._field_shape(which))
        self.lib.ref_set_field(self.h, ic, which, arr.reshape(-1))

    def get_field(self, ic, which):
        out = np.zeros(self._field_shape(which), dtype=np.float64)
        self.lib.ref_get_field(self.h, ic, which, out.reshape(-1))
        return out

    def set_particles(self, ic, isp, xu, np_alloc=None):
        xu = np.ascontiguousarray(xu, dtype=np.float64).reshape(-1, 7)
        n = xu.shape[0]
        if np_alloc is None:
            np_alloc = int(n * (1 + self.cfg.buffer_ratio))
        self.lib.ref_set_particles(self.h, ic, isp, xu.reshape(-1), n, np_alloc)

    def get_np(self, ic, isp):
        return self.lib.ref_get_np(self.h, ic, isp)

    def get_np_total(self, ic, isp):
        return self.lib.ref_get_np_total(self.h, ic, isp)

    def get_particles(self, ic, isp, which=0, n=None):
        if n is None:
            n = self.get_np(ic, isp)
        out = np.zeros((n, 7), dtype=np.float64)
        if n > 0:
            self.lib.ref_get_particles(self.h, ic, isp, which, n, out.reshape(-1))
        return out

    def get_pindex(self, ic, isp):
        out = np.zeros(self.Ng + 1, dtype=np.int32)
        self.lib.ref_get_pindex(self.h, ic, isp, out)
        return out

    def get_gindex(self, ic, isp, n=None):
        if n is None:
            n = self.get_np(ic, isp)
        out = np.zeros(max(n, 1), dtype=np.int32)
        if n > 0:
            self.lib.ref_get_gindex(self.h, ic, isp, n, out)
        return out[:n]

    # -- phases ----------------------------------------------------------------------------
    def finalize_setup(self):
        self.lib.ref_finalize_setup(self.h)

    def init_friedman(self):
        self.lib.ref_init_friedman(self.h)

    def push_bfd(self, dt):
        self.lib.ref_push_bfd(self.h, dt)

    def push_efd(self, dt):
        self.lib.ref_push_efd(self.h, dt)

    def push_velocity(self, dt):
        self.lib.ref_push_velocity(self.h, dt)

    def push_position(self, dt):
        self.lib.ref_push_position(self.h, dt)

    def deposit_current(self, dt):
        self.lib.ref_deposit_current(self.h, dt)

    def deposit_moment(self):
        self.lib.ref_deposit_moment(self.h)

    def sort_particle(self):
        self.lib.ref_sort_particle(self.h)

    def exchange(self, mode):
        self.lib.ref_exchange(self.h, mode)

    def step(self, dt, nstep=1):
        self.lib.ref_step(self.h, dt, nstep)

    def get_diverror(self):
        e, b = C.c_double(), C.c_double()
        out = np.zeros((self.nchunk, 2))
        for ic in range(self.nchunk):
            self.lib.ref_get_diverror(self.h, ic, C.byref(e), C.byref(b))
            out[ic] = e.value, b.value
        return out

    def get_energy(self):
        e, b = C.c_double(), C.c_double()
        out = np.zeros((self.nchunk, 2 + self.Ns))
        p = np.zeros(self.Ns)
        for ic in range(self.nchunk):
            self.lib.ref_get_energy(self.h, ic, C.byref(e), C.byref(b), p)
            out[ic, 0], out[ic, 1] = e.value, b.value
            out[ic, 2:] = p
        return out
