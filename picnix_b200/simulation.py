"""Host-side Python mirror of the reference's chunk-level API on top of the C ABI.

`CudaSim` holds one device arena (all chunks of one rank) and exposes the entry points of
`PicChunk` (pic/pic_chunk.hpp:90-143) batched over the rank's chunks, plus the step schedule of
`PicApplication::push_openmp` (pic/pic_application.cpp:219-292).  Method names and argument meaning
follow the reference so that parity tests read like tests of the reference.

Everything numerical happens inside libpicnix_b200.so (CUDA, sm_100a).  This module is plumbing.
"""
import ctypes as C

import numpy as np

from . import capi


class CudaSim:
    name = "cuda"

    def __init__(self, ndims, cdims, Ns, cc, delh=1.0, order=2, pusher=0, interp=0, periodic=(1, 1, 1),
                 friedman=0.0, buffer_ratio=0.2, nrank=1, rank=0, boundary=None):
        self.lib = capi.load()
        cfg = capi.Config()
        cfg.ndims[:] = ndims
        cfg.cdims[:] = cdims
        cfg.periodic[:] = periodic
        cfg.order, cfg.pusher, cfg.interp, cfg.Ns = order, pusher, interp, Ns
        cfg.nrank, cfg.rank = nrank, rank
        cfg.cc = cc
        if np.isscalar(delh):
            cfg.delx = cfg.dely = cfg.delz = delh
        else:
            cfg.delz, cfg.dely, cfg.delx = delh
        cfg.friedman, cfg.buffer_ratio = friedman, buffer_ratio
        self.cfg = cfg
        self.Ns = Ns
        self.h = C.c_void_p()
        bptr = None
        if boundary is not None:
            self._boundary = np.ascontiguousarray(boundary, dtype=np.int32)
            bptr = self._boundary.ctypes.data_as(C.c_void_p)
        status = self.lib.picnix_cuda_arena_create(C.byref(cfg), bptr, C.byref(self.h))
        if status != capi.OK:
            msg = self.lib.picnix_cuda_last_error(self.h).decode() if self.h else "arena_create failed"
            if self.h:
                self.lib.picnix_cuda_arena_destroy(self.h)
                self.h = C.c_void_p()
            raise capi.PicnixError(status, msg)
        nchunk, begin, margin, Ng = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        dims = np.zeros(3, dtype=np.int32)
        self._check(self.lib.picnix_cuda_get_layout(self.h, C.byref(nchunk), C.byref(begin), dims,
                                                    C.byref(margin), C.byref(Ng)))
        self.nchunk = nchunk.value
        self.chunk_id_begin = begin.value
        self.shape = tuple(int(v) for v in dims)
        self.nb = margin.value
        self.Ng = Ng.value
        self._pending = {}
        self._capacity_set = False
        self._species = {}
        self._stream_ptr = None  # the arena owns its stream until set_stream() is called
        self._options = {}       # set_option history: carried over when the arena is rebuilt (rebalance)
        self._bcs = []           # set_boundary_condition history, likewise
        self._ctor = dict(ndims=tuple(ndims), cdims=tuple(cdims), Ns=Ns, cc=cc, delh=delh, order=order,
                          pusher=pusher, interp=interp, periodic=tuple(periodic), friedman=friedman,
                          buffer_ratio=buffer_ratio)

    # -- plumbing --------------------------------------------------------------------------
    def _check(self, status):
        if status != capi.OK:
            raise capi.PicnixError(status, self.lib.picnix_cuda_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.picnix_cuda_arena_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self._check(self.lib.picnix_cuda_synchronize(self.h))

    def set_option(self, key, value):
        self._check(self.lib.picnix_cuda_set_option(self.h, key.encode(), int(value)))
        self._options[key] = int(value)

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.picnix_cuda_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))
        self._stream_ptr = int(cuda_stream_ptr)

    def counters(self):
        launches, pushes = C.c_int64(), C.c_int64()
        self._check(self.lib.picnix_cuda_get_counters(self.h, C.byref(launches), C.byref(pushes)))
        return launches.value, pushes.value

    def set_boundary_condition(self, axis, side, kind, values=None):
        """Physical boundary of one face (axis 0 = z, 1 = y, 2 = x; side 0 = lower, 1 = upper): the device
        version of the set_boundary_field / set_boundary_particle hooks of the reference's examples."""
        vals = None
        if values is not None:
            self._bc_values = np.ascontiguousarray(values, dtype=np.float64)
            vals = self._bc_values.ctypes.data_as(C.c_void_p)
        self._check(self.lib.picnix_cuda_set_boundary_condition(self.h, axis, side, kind, vals))
        self._bcs.append((axis, side, kind, None if values is None else list(values)))

    def inject_particles(self, ic, isp, xu):
        """PicChunk::inject_particle: append host-generated particles (between boundary_begin and
        boundary_end of the particle exchange)."""
        self.commit()
        xu = np.ascontiguousarray(xu, dtype=np.float64).reshape(-1, 7)
        self._check(self.lib.picnix_cuda_inject_particles(self.h, ic, isp, xu.reshape(-1), xu.shape[0]))

    def growth_stats(self):
        """(segment re-layouts, migrants delivered one step late) since the arena was created."""
        regrows, late = C.c_int64(), C.c_int64()
        self._check(self.lib.picnix_cuda_get_growth_stats(self.h, C.byref(regrows), C.byref(late)))
        return regrows.value, late.value

    # -- decomposition ---------------------------------------------------------------------
    def chunkmap(self):
        return capi.sfc_build(*tuple(self.cfg.cdims))

    def neighbors(self, ic):
        nbid = np.zeros(27, dtype=np.int32)
        nbrank = np.zeros(27, dtype=np.int32)
        self._check(self.lib.picnix_cuda_get_neighbors(self.h, ic, nbid, nbrank))
        return nbid, nbrank

    # -- state -----------------------------------------------------------------------------
    def _field_shape(self, which):
        tail = {capi.FIELD_UF: (6,), capi.FIELD_UJ: (4,), capi.FIELD_FF: (3, 6), capi.FIELD_UM: (self.Ns, 14)}[which]
        return self.shape + tail

    def set_species(self, isp, q, m):
        self._species[isp] = (q, m)
        self._check(self.lib.picnix_cuda_set_species(self.h, isp, q, m))

    def set_capacity(self, caps):
        """XtensorParticle::allocate for every (chunk, species) segment at once."""
        caps = np.ascontiguousarray(caps, dtype=np.int32).reshape(-1)
        assert caps.size == self.nchunk * self.Ns
        self._check(self.lib.picnix_cuda_set_particle_capacity(self.h, caps))
        self._capacity_set = True

    # -- chunk moves (rebalancing): the whole state of a local chunk as one device buffer --------
    def chunk_pack_size(self, ic):
        self.commit()
        n = C.c_int64()
        self._check(self.lib.picnix_cuda_chunk_pack_size(self.h, ic, C.byref(n)))
        return n.value

    def chunk_pack(self, ic):
        import torch

        self.commit()
        n = C.c_int64()
        self._check(self.lib.picnix_cuda_chunk_pack_size(self.h, ic, C.byref(n)))
        buf = torch.empty(n.value, dtype=torch.uint8, device="cuda")
        self._check(self.lib.picnix_cuda_chunk_pack(self.h, ic, C.c_void_p(buf.data_ptr()), n.value))
        return buf

    def chunk_unpack(self, ic, buf):
        self._check(self.lib.picnix_cuda_chunk_unpack(self.h, ic, C.c_void_p(buf.data_ptr()), buf.numel()))

    def set_field(self, ic, which, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        assert arr.shape == self._field_shape(which), (arr.shape, self._field_shape(which))
        self._check(self.lib.picnix_cuda_upload_field(self.h, ic, which, arr.reshape(-1)))

    def get_field(self, ic, which):
        out = np.zeros(self._field_shape(which), dtype=np.float64)
        self._check(self.lib.picnix_cuda_download_field(self.h, ic, which, out.reshape(-1)))
        return out

    def set_particles(self, ic, isp, xu, np_alloc=None):
        """Stage the particles of (chunk, species); they are uploaded by commit()."""
        xu = np.ascontiguousarray(xu, dtype=np.float64).reshape(-1, 7)
        if np_alloc is None:
            np_alloc = int(xu.shape[0] * (1 + self.cfg.buffer_ratio))
        if self._capacity_set:
            # capacities are fixed: upload directly
            self._check(self.lib.picnix_cuda_upload_particles(self.h, ic, isp, xu.reshape(-1), xu.shape[0]))
        else:
            self._pending[(ic, isp)] = (xu, np_alloc)

    def commit(self):
        """Fix the segment capacities (XtensorParticle::allocate) and upload staged particles."""
        if self._capacity_set:
            return
        caps = np.zeros(self.nchunk * self.Ns, dtype=np.int32)
        for (ic, isp), (xu, np_alloc) in self._pending.items():
            caps[ic * self.Ns + isp] = np_alloc
        self._check(self.lib.picnix_cuda_set_particle_capacity(self.h, caps))
        self._capacity_set = True
        for (ic, isp), (xu, _) in self._pending.items():
            self._check(self.lib.picnix_cuda_upload_particles(self.h, ic, isp, xu.reshape(-1), xu.shape[0]))
        self._pending = {}

    def get_np_all(self):
        self.commit()
        out = np.zeros(self.nchunk * self.Ns, dtype=np.int32)
        self._check(self.lib.picnix_cuda_get_np(self.h, out))
        return out.reshape(self.nchunk, self.Ns)

    def get_np(self, ic, isp):
        return int(self.get_np_all()[ic, isp])

    def get_particles(self, ic, isp, which=0, n=None):
        self.commit()
        if n is None:
            n = self.get_np(ic, isp)
        out = np.zeros((n, 7), dtype=np.float64)
        if n > 0:
            self._check(self.lib.picnix_cuda_download_particles(self.h, ic, isp, which, n, out.reshape(-1)))
        return out

    def get_pindex(self, ic, isp):
        self.commit()
        out = np.zeros(self.Ng + 1, dtype=np.int32)
        self._check(self.lib.picnix_cuda_download_pindex(self.h, ic, isp, out))
        return out

    def get_gindex(self, ic, isp, n=None):
        self.commit()
        if n is None:
            n = self.get_np(ic, isp)
        out = np.zeros(max(n, 1), dtype=np.int32)
        if n > 0:
            self._check(self.lib.picnix_cuda_download_gindex(self.h, ic, isp, n, out))
        return out[:n]

    # -- phases (all local chunks at once) -------------------------------------------------
    def finalize_setup(self):
        """Tail of MainChunk::setup + PicApplication::setup_chunks (pic/pic_application.cpp:106-130)."""
        self.commit()
        self.init_friedman()
        self.sort_particle()
        self.exchange(capi.BOUNDARY_EMF)

    def init_friedman(self):
        self._check(self.lib.picnix_cuda_init_friedman(self.h, 0, -1))

    def push_bfd(self, dt):
        self._check(self.lib.picnix_cuda_push_bfd(self.h, 0, -1, dt))

    def push_efd(self, dt):
        self._check(self.lib.picnix_cuda_push_efd(self.h, 0, -1, dt))

    def push_velocity(self, dt):
        self.commit()
        self._check(self.lib.picnix_cuda_push_velocity(self.h, 0, -1, dt))

    def push_position(self, dt):
        self.commit()
        self._check(self.lib.picnix_cuda_push_position(self.h, 0, -1, dt))

    def deposit_current(self, dt):
        self.commit()
        self._check(self.lib.picnix_cuda_deposit_current(self.h, 0, -1, dt))

    def deposit_moment(self):
        self.commit()
        self._check(self.lib.picnix_cuda_deposit_moment(self.h))

    def get_energy(self):
        """PicChunk::get_energy per chunk: columns efd, bfd, particle[0..Ns-1] (after deposit_moment
        and the BoundaryMom exchange, like the reference's history diagnostic)."""
        fe = self.get_field_energy()
        p = np.zeros(self.nchunk * self.Ns)
        self._check(self.lib.picnix_cuda_get_particle_energy(self.h, p))
        return np.concatenate([fe, p.reshape(self.nchunk, self.Ns)], axis=1)

    def push_deposit_fused(self, dt):
        self.commit()
        self._check(self.lib.picnix_cuda_push_deposit_fused(self.h, 0, -1, dt))

    def sort_particle(self):
        self.commit()
        self._check(self.lib.picnix_cuda_sort_particle(self.h, 0, -1))

    def boundary_begin(self, mode):
        self.commit()
        self._check(self.lib.picnix_cuda_boundary_begin(self.h, mode))

    def boundary_end(self, mode):
        self._check(self.lib.picnix_cuda_boundary_end(self.h, mode))

    def exchange(self, mode):
        """pack + begin + (probe) + end + unpack for all chunks; single-rank arenas only."""
        if self.cfg.nrank != 1:
            raise RuntimeError("exchange() on a multi-rank arena needs a transport (see distributed.py)")
        self.boundary_begin(mode)
        self.boundary_end(mode)

    def step(self, dt, nstep=1):
        self.commit()
        self._check(self.lib.picnix_cuda_step(self.h, dt, nstep))

    # -- host-buffer step (what a host-resident PicChunk would call) --------------------------
    def host_state(self, pinned=True):
        """Current device state as HOST arrays in the reference's layouts (uf, uj, ff[..][3][6], AoS
        particles with `caps[seg]` slots per (chunk, species) segment) -- the argument of step_host."""
        self.commit()
        Ns, nchunk, ncell = self.Ns, self.nchunk, self.Ng

        def alloc(n):
            if pinned:
                import torch

                return torch.zeros(int(n), dtype=torch.float64, pin_memory=True).numpy()
            return np.zeros(int(n), dtype=np.float64)

        np_now = self.get_np_all().reshape(-1).astype(np.int32)
        caps = np.array([int(n * (1 + self.cfg.buffer_ratio)) for n in np_now], dtype=np.int32)
        caps = ((caps + 128) // 128) * 128
        st = {"uf": alloc(nchunk * ncell * 6), "uj": alloc(nchunk * ncell * 4), "ff": alloc(nchunk * ncell * 18),
              "xu": alloc(int(caps.sum()) * 7), "np": np_now.copy(), "caps": caps}
        off = 0
        for ic in range(nchunk):
            st["uf"][ic * ncell * 6:(ic + 1) * ncell * 6] = self.get_field(ic, capi.FIELD_UF).reshape(-1)
            st["uj"][ic * ncell * 4:(ic + 1) * ncell * 4] = self.get_field(ic, capi.FIELD_UJ).reshape(-1)
            st["ff"][ic * ncell * 18:(ic + 1) * ncell * 18] = self.get_field(ic, capi.FIELD_FF).reshape(-1)
            for isp in range(Ns):
                seg = ic * Ns + isp
                n = int(np_now[seg])
                st["xu"][off * 7:(off + n) * 7] = self.get_particles(ic, isp, 0, n).reshape(-1)
                off += int(caps[seg])
        return st

    def step_host(self, st, dt, nstep=1):
        """picnix_cuda_step_host: host arrays in, `nstep` steps on the device, host arrays out."""
        np_out = np.zeros_like(st["np"])
        self._check(self.lib.picnix_cuda_step_host(self.h, dt, nstep, st["uf"], st["uj"], st["ff"], st["xu"],
                                                   st["np"], st["caps"], np_out))
        st["np"] = np_out
        return st

    def upload_state(self, st):
        """picnix_cuda_upload_state: whole-rank host arrays -> device (particles end up cell-ordered)."""
        self.commit()
        self._check(self.lib.picnix_cuda_upload_state(self.h, st["uf"], st["uj"], st["ff"], st["xu"], st["np"],
                                                      st["caps"]))

    def download_state(self, st):
        """picnix_cuda_download_state: device -> the same host arrays; st['np'] gets the new counts."""
        np_out = np.zeros_like(st["np"])
        self._check(self.lib.picnix_cuda_download_state(self.h, st["uf"], st["uj"], st["ff"], st["xu"],
                                                        st["caps"], np_out))
        st["np"] = np_out
        return st

    @staticmethod
    def host_particles(st, Ns, ic, isp):
        seg = ic * Ns + isp
        off = int(st["caps"][:seg].sum())
        return st["xu"][off * 7:(off + int(st["np"][seg])) * 7].reshape(-1, 7)

    def get_diverror(self):
        e = np.zeros(self.nchunk)
        b = np.zeros(self.nchunk)
        self._check(self.lib.picnix_cuda_get_diverror(self.h, e, b))
        return np.stack([e, b], axis=1)

    def get_field_energy(self):
        e = np.zeros(self.nchunk)
        b = np.zeros(self.nchunk)
        self._check(self.lib.picnix_cuda_get_field_energy(self.h, e, b))
        return np.stack([e, b], axis=1)

    # -- multi-rank plumbing ---------------------------------------------------------------
    def peers(self):
        n = C.c_int32()
        self._check(self.lib.picnix_cuda_get_peers(self.h, C.byref(n), None))
        ranks = np.zeros(max(n.value, 1), dtype=np.int32)
        self._check(self.lib.picnix_cuda_get_peers(self.h, C.byref(n), ranks.ctypes.data_as(C.c_void_p)))
        return [int(r) for r in ranks[: n.value]]

    def comm_buffer(self, mode, peer_index):
        sp, rp = C.c_void_p(), C.c_void_p()
        sb, rb = C.c_int64(), C.c_int64()
        self._check(self.lib.picnix_cuda_get_comm_buffer(self.h, mode, peer_index, C.byref(sp), C.byref(sb),
                                                         C.byref(rp), C.byref(rb)))
        return sp.value, sb.value, rp.value, rb.value

    def set_recv_bytes(self, mode, peer_index, nbytes):
        self._check(self.lib.picnix_cuda_set_recv_bytes(self.h, mode, peer_index, nbytes))
