"""One-process-per-GPU driver: the reference's MPI rank decomposition mapped onto torch.distributed.

The reference gives every MPI rank a contiguous range of space-filling-curve chunk ids
(nix/application.cpp:245-309, nix/balancer.cpp:101-124) and exchanges three halos per step between
neighbouring chunks (pic/pic_application.cpp:219-292).  Here every rank owns one device arena; halos
between chunks of the same arena never leave the GPU, and the chunks owned by another rank are
served by ONE aggregated send and ONE receive buffer per peer and mode, moved with NCCL send/recv
(batched in a single group) over NVLink.  torch.distributed is plumbing only: rendezvous, the
send/recv of the buffers the CUDA library packed, and the final timing reduction.

`Transport` is backend agnostic (device buffers for the CUDA arena, host buffers for the CPU oracle
that the gloo tests use), so the message planning is covered by world_size-2 CPU tests.
"""
import os
import time

import numpy as np

from . import capi
from .simulation import CudaSim

MODE_EMF, MODE_CUR, MODE_MOM, MODE_PARTICLE = 0, 1, 2, 3


class _CudaBuffer:
    """Zero-copy view of a raw device pointer for torch.as_tensor."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def cuda_view(ptr, nbytes):
    import torch

    if nbytes == 0:
        return torch.empty(0, dtype=torch.uint8, device="cuda")
    return torch.as_tensor(_CudaBuffer(ptr, nbytes), device="cuda")


def host_view(ptr, nbytes):
    import ctypes

    import torch

    if nbytes == 0:
        return torch.empty(0, dtype=torch.uint8)
    arr = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(nbytes,))
    return torch.from_numpy(arr)


class Transport:
    """Moves the per-peer halo buffers of `sim` between ranks with torch.distributed."""

    def __init__(self, sim, world, device_buffers=True):
        self.sim = sim
        self.world = world
        self.view = cuda_view if device_buffers else host_view
        self.device = "cuda" if device_buffers else "cpu"
        self.peers = sim.peers() if world > 1 else []
        # transfers overlap the kernels enqueued between start() and finish(); PICNIX_OVERLAP=0 restores
        # the strictly serial pack -> transfer -> unpack of round 1
        self.overlap = os.environ.get("PICNIX_OVERLAP", "1") != "0" and device_buffers

    def _shares_stream(self):
        """True when the arena enqueues on the stream torch (and therefore NCCL) orders against."""
        if self.device != "cuda":
            return True
        import torch

        return getattr(self.sim, "_stream_ptr", None) == torch.cuda.current_stream().cuda_stream

    def _issue(self, pairs):
        """Enqueue one group of sends / receives.  NCCL runs on its own stream, ordered after everything
        enqueued on the current stream so far (the pack kernels); kernels enqueued afterwards overlap the
        transfer until _complete() makes the current stream wait for it."""
        import torch.distributed as dist

        if not self._shares_stream():
            # the arena packs on its own stream: NCCL must not read the buffers before that is done
            self.sim.synchronize()
        ops = []
        for peer, send, recv in pairs:
            if recv.numel() > 0:
                ops.append(dist.P2POp(dist.irecv, recv, peer))
            if send.numel() > 0:
                ops.append(dist.P2POp(dist.isend, send, peer))
        return dist.batch_isend_irecv(ops) if ops else []

    def _complete(self, reqs):
        for req in reqs:
            req.wait()
        if reqs and not self._shares_stream():
            # ... and the arena must not unpack before the data has landed
            import torch

            torch.cuda.current_stream().synchronize()

    def _exchange(self, pairs):
        self._complete(self._issue(pairs))

    def start(self, mode):
        """Non-blocking half of move(): after boundary_begin(mode), start the transfer and return a
        ticket for finish().  The reference does the same with MPI_Isend / MPI_Irecv in
        Chunk::begin_bc_exchange (nix/chunk.hpp:464-524) and goes on computing
        (pic/pic_application.cpp:242-251)."""
        if not self.peers:
            return []
        if mode == MODE_PARTICLE:
            bufs = [self.sim.comm_buffer(mode, i) for i in range(len(self.peers))]
            if any(b[3] == 0 for b in bufs):
                self.move(mode)  # synchronous count exchange (first steps): nothing left to overlap
                return []
        pairs = []
        for i, peer in enumerate(self.peers):
            sp, sb, rp, rb = self.sim.comm_buffer(mode, i)
            pairs.append((peer, self.view(sp, sb), self.view(rp, rb)))
        return self._issue(pairs)

    def finish(self, ticket):
        """Before boundary_end(mode): the current stream waits for the transfer started by start()."""
        self._complete(ticket)

    def move(self, mode):
        """Between boundary_begin(mode) and boundary_end(mode): send -> peer's recv buffer."""
        if not self.peers:
            return
        import torch

        if mode == MODE_PARTICLE:
            bufs = [self.sim.comm_buffer(mode, i) for i in range(len(self.peers))]
            if any(b[3] == 0 for b in bufs):
                # synchronous protocol, variable size: first the byte counts (the reference sizes its
                # receive with MPI_Iprobe + MPI_Get_count, nix/chunk.cpp:329-345), then the payload.
                # With the lagged-count protocol (option async_migration) both sizes are already known.
                scount = [torch.tensor([b[1]], dtype=torch.int64, device=self.device) for b in bufs]
                rcount = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in bufs]
                self._exchange([(p, s, r) for p, s, r in zip(self.peers, scount, rcount)])
                counts = torch.cat(rcount).cpu().tolist()  # one synchronisation for all peers
                for i, nbytes in enumerate(counts):
                    self.sim.set_recv_bytes(mode, i, int(nbytes))
        pairs = []
        for i, peer in enumerate(self.peers):
            sp, sb, rp, rb = self.sim.comm_buffer(mode, i)
            pairs.append((peer, self.view(sp, sb), self.view(rp, rb)))
        self._exchange(pairs)


def step_phases(sim, transport, dt, kernel_events=None):
    """One time step in the order of PicApplication::push_openmp (pic/pic_application.cpp:219-292).

    `sim` is any backend with the chunk-level API (the CUDA arena, or the CPU oracle in the gloo
    tests); `transport.move(mode)` carries the per-peer buffers between boundary_begin and
    boundary_end.  Phase A: B half step, push + deposit, start the J and particle exchanges, B half
    step; phase B: finish J, E step, start the E/B exchange; phases C-E: particles arrive (wrap,
    count, sort), fields arrive.
    """
    overlap = hasattr(transport, "start") and getattr(transport, "overlap", True)
    start = transport.start if overlap else (lambda mode: transport.move(mode))
    finish = transport.finish if overlap else (lambda ticket: None)

    sim.push_bfd(0.5 * dt)
    if kernel_events is not None:
        kernel_events[0].record()
    sim.push_deposit_fused(dt)
    if kernel_events is not None:
        kernel_events[1].record()
    # the J and particle transfers run under the second B half step, the J unpack and the E step; the
    # E/B transfer under the particle unpack and the sort (pic/pic_application.cpp:242-251)
    sim.boundary_begin(MODE_CUR)
    t_cur = start(MODE_CUR)
    sim.boundary_begin(MODE_PARTICLE)
    t_par = start(MODE_PARTICLE)
    sim.push_bfd(0.5 * dt)
    finish(t_cur)
    sim.boundary_end(MODE_CUR)
    sim.push_efd(dt)
    sim.boundary_begin(MODE_EMF)
    t_emf = start(MODE_EMF)
    finish(t_par)
    sim.boundary_end(MODE_PARTICLE)
    finish(t_emf)
    sim.boundary_end(MODE_EMF)


def exchange(sim, transport, mode):
    sim.boundary_begin(mode)
    transport.move(mode)
    sim.boundary_end(mode)


def _owner(boundary, gid):
    return int(np.searchsorted(np.asarray(boundary), gid, side="right") - 1)


def _capacities(np_rows, ratio):
    """segment capacities for incoming chunks: the reference's (1 + buffer_ratio) slack plus room
    for one allocation unit, so that a freshly moved chunk can still receive migrants"""
    return (np_rows.astype(np.float64) * (1 + ratio)).astype(np.int32).reshape(-1) + 256


def rebalance_in_process(sims, new_boundary, make_sim):
    """Application::rebalance for R arenas living in ONE process (tests; single-node tools).

    `sims[r]` owns the chunk ids [b[r], b[r+1]); `new_boundary` comes from Balancer::assign
    (capi.assign_rebalance / assign_initial).  Every chunk is packed on its old owner
    (PicChunk::pack -> picnix_cuda_chunk_pack), handed over as a device buffer and unpacked on its
    new owner; new arenas are created by `make_sim(rank, boundary)`.  Returns the new arenas."""
    nrank = len(sims)
    old_boundary = [s.chunk_id_begin for s in sims] + [sims[-1].chunk_id_begin + sims[-1].nchunk]
    np_all = np.concatenate([s.get_np_all() for s in sims], axis=0)
    new = [make_sim(r, new_boundary) for r in range(nrank)]
    for r, t in enumerate(new):
        t.set_capacity(_capacities(np_all[new_boundary[r]:new_boundary[r + 1]], t.cfg.buffer_ratio))
        for isp, (q, m) in sims[0]._species.items():
            t.set_species(isp, q, m)
    for gid in range(old_boundary[-1]):
        src, dst = _owner(old_boundary, gid), _owner(new_boundary, gid)
        buf = sims[src].chunk_pack(gid - old_boundary[src])
        new[dst].chunk_unpack(gid - new_boundary[dst], buf)
    for t in new:
        t.sort_particle()  # keys + pindex of the arrived particles
    return new


class DistributedSim(CudaSim):
    """CudaSim whose chunk ids are split over `world` ranks like the reference's MPI ranks."""

    def __init__(self, ndims, cdims, Ns, cc, rank=0, world=1, boundary=None, async_migration=True, **kw):
        super().__init__(ndims, cdims, Ns, cc, nrank=world, rank=rank, boundary=boundary, **kw)
        self.rank, self.world = rank, world
        self._async_migration = async_migration
        if os.environ.get("PICNIX_ASYNC_MIGRATION", "1") == "0":
            async_migration = False  # escape hatch: the synchronous count exchange of the first version
        if world > 1 and async_migration:
            # particle exchange without a host synchronisation in the middle of the step (halo.cu)
            self.set_option("async_migration", 1)
        self.transport = Transport(self, world, device_buffers=True)

    def exchange(self, mode):
        self.boundary_begin(mode)
        self.transport.move(mode)
        self.boundary_end(mode)

    def step_phases(self, dt, kernel_events=None):
        """One time step, PicApplication::push_openmp order, transports between begin and end."""
        self.commit()
        step_phases(self, self.transport, dt, kernel_events)

    def step(self, dt, nstep=1):
        if self.world == 1:
            return super().step(dt, nstep)
        for _ in range(nstep):
            self.step_phases(dt)

    def rebalanced(self, new_boundary):
        """Application::rebalance across GPUs: returns a NEW DistributedSim that owns the chunk range
        `new_boundary[rank] .. new_boundary[rank+1]`; chunks that change owner travel packed
        (picnix_cuda_chunk_pack) over NCCL send/recv, device to device."""
        import torch
        import torch.distributed as dist

        new_boundary = [int(b) for b in new_boundary]
        mine = torch.zeros((self.cfg.cdims[0] * self.cfg.cdims[1] * self.cfg.cdims[2], self.Ns), dtype=torch.int64,
                           device="cuda")
        mine[self.chunk_id_begin:self.chunk_id_begin + self.nchunk] = torch.from_numpy(
            self.get_np_all().astype(np.int64)).cuda()
        if self.world > 1:
            dist.all_reduce(mine)  # every chunk has exactly one owner: sum == gather
        np_all = mine.cpu().numpy()
        old_boundary = [0] * (self.world + 1)
        ob = torch.zeros(self.world + 1, dtype=torch.int64, device="cuda")
        ob[self.rank] = self.chunk_id_begin
        if self.rank == self.world - 1:
            ob[self.world] = self.chunk_id_begin + self.nchunk
        if self.world > 1:
            dist.all_reduce(ob)
        old_boundary = [int(v) for v in ob.cpu()]

        new = DistributedSim(rank=self.rank, world=self.world, boundary=new_boundary,
                             async_migration=self._async_migration, **self._ctor)
        # the new arena keeps enqueuing on the caller's stream and keeps the tuning switches
        if self._stream_ptr is not None:
            new.set_stream(self._stream_ptr)
        for key, value in self._options.items():
            if key != "async_migration":
                new.set_option(key, value)
        for axis, side, kind, values in self._bcs:
            new.set_boundary_condition(axis, side, kind, values)
        new.set_capacity(_capacities(np_all[new_boundary[self.rank]:new_boundary[self.rank + 1]],
                                     self.cfg.buffer_ratio))
        for isp, (q, m) in self._species.items():
            new.set_species(isp, q, m)
        # chunks that stay are handed over device to device; the ones that change owner travel in ONE
        # group of sends and receives: the packed sizes of all chunks are agreed on first (one reduction,
        # one host read), so no message needs a size handshake of its own
        nchunk_all = old_boundary[-1]
        moves = [(gid, _owner(old_boundary, gid), _owner(new_boundary, gid)) for gid in range(nchunk_all)]
        sizes = torch.zeros(nchunk_all, dtype=torch.int64, device="cuda")
        mine_sizes = {}
        for gid, src, dst in moves:
            if src == self.rank and dst != self.rank:
                mine_sizes[gid] = self.chunk_pack_size(gid - old_boundary[src])
        if mine_sizes:
            idx = torch.tensor(list(mine_sizes.keys()), dtype=torch.int64, device="cuda")
            sizes[idx] = torch.tensor(list(mine_sizes.values()), dtype=torch.int64, device="cuda")
        if self.world > 1:
            dist.all_reduce(sizes)
        sizes = sizes.cpu().numpy()
        ops, arrivals, keep = [], [], []
        for gid, src, dst in moves:
            if src == self.rank and dst == self.rank:
                new.chunk_unpack(gid - new_boundary[dst], self.chunk_pack(gid - old_boundary[src]))
            elif src == self.rank:
                buf = self.chunk_pack(gid - old_boundary[src])
                keep.append(buf)
                ops.append(dist.P2POp(dist.isend, buf, dst))
            elif dst == self.rank:
                buf = torch.empty(int(sizes[gid]), dtype=torch.uint8, device="cuda")
                arrivals.append((gid, buf))
                ops.append(dist.P2POp(dist.irecv, buf, src))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for gid, buf in arrivals:
            new.chunk_unpack(gid - new_boundary[self.rank], buf)
        new.sort_particle()
        return new

    # -- end-to-end measurement through the host-buffer entry point -----------------------------
    def measure_e2e(self, dt, nstep, barrier=None):
        """Time `nstep` steps with the rank's whole state crossing the host boundary every step.

        One rank: picnix_cuda_step_host.  Several ranks: picnix_cuda_upload_state, the phases with
        the NCCL transport between boundary_begin/end, picnix_cuda_download_state -- the same
        pipeline, with the inter-GPU halos in the middle.  Returns this rank's wall time and byte
        counts; the caller takes the max / sums over ranks."""
        import torch

        st = self.host_state(pinned=True)  # outside the timed region
        np_now = st["np"].copy()
        fields = st["uf"].nbytes + st["uj"].nbytes + st["ff"].nbytes

        def one_step():
            if self.world == 1:
                self.step_host(st, dt, 1)
                return fields - st["uj"].nbytes  # uj is an output only
            self.upload_state(st)
            self.step_phases(dt)
            self.download_state(st)
            return fields

        one_step()  # untimed: allocates the slabs and streams of the pipeline
        h2d = d2h = 0
        torch.cuda.synchronize()
        if barrier is not None:
            barrier()
        t0 = time.perf_counter()
        for _ in range(nstep):
            n_in = int(st["np"].sum())
            h2d += one_step() + n_in * 56
            d2h += fields + int(st["np"].sum()) * 56
        torch.cuda.synchronize()
        if barrier is not None:
            barrier()
        elapsed = time.perf_counter() - t0
        return {"elapsed": elapsed, "particles": float(np_now.sum()), "h2d": h2d // nstep, "d2h": d2h // nstep,
                "steps": nstep}
