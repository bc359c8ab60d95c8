// -*- C++ -*-
// Host-buffer step: what a host-resident PicChunk sees when it hands its arrays to the GPU for a
// time step (picnix_cuda_step_host, include/picnix_b200.h).
//
// The reference keeps uf/uj/ff and the AoS particle array [Np][7] of every chunk on the host
// (pic/pic_chunk.cpp:106-122, nix/xtensor_particle.hpp:49-65).  One call moves that state to the
// device, runs the step schedule of PicApplication::push_openmp on it and moves it back.  The call
// is PCIe-bound (T3D: 8.4 GB each way against 42 ms of kernels), so it is organised as a pipeline:
//
//   * host buffers are used in place.  Pageable buffers are page-locked once (cudaHostRegister)
//     and remembered, because the reference's arrays live as long as the chunk does;
//   * three streams: copy-in, the arena's compute stream, copy-out.  Particles travel in batches
//     of whole (chunk, species) segments through a ring of device slabs: while slab k is on the
//     wire, the AoS<->SoA transpose of slab k-1 runs on the compute stream;
//   * uf and uj have the device layout already and are copied straight into place; ff
//     ([cell][3][6] on the host, [cell][3][3] on the device) goes through the slabs;
//   * uj is not uploaded when nstep >= 1: deposit_current starts with fill_all(uj, 0)
//     (pic/engine/current.hpp:91,152), its previous content cannot influence the result;
//   * uploaded particles get their cell keys and pindex from one count + sort on the device, so the
//     tiled push+deposit kernel (which needs cell-ordered particles) applies to the first step.
#include "arena.hpp"
#include "transpose_kernels.cuh"

#include <algorithm>
#include <cstring>

namespace picnix
{

struct HostIO {
  static constexpr int NSLOT = 3;
  cudaStream_t         h2d = nullptr, d2h = nullptr;
  double*              slab[NSLOT] = {nullptr, nullptr, nullptr};
  int64_t              slab_elems  = 0;
  cudaEvent_t          filled[NSLOT], drained[NSLOT];
  cudaEvent_t          ev_misc     = nullptr;
  bool                 used[NSLOT] = {false, false, false};
  bool                 ready       = false;
  std::vector<void*>   registered; // buffers this library page-locked
  int64_t*             d_capoff    = nullptr; // prefix sums of the host capacities (grouped transfers)
  int64_t*             h_capoff    = nullptr; // pinned staging of the same
  int                  capoff_len  = 0;
};

void hostio_destroy(picnix_arena* a)
{
  HostIO* io = a->hostio;
  if (io == nullptr)
    return;
  for (void* p : io->registered)
    cudaHostUnregister(p); // the owner may have freed it already; errors are irrelevant here
  cudaGetLastError();
  for (int s = 0; s < HostIO::NSLOT; s++) {
    if (io->slab[s])
      cudaFree(io->slab[s]);
    if (io->ready) {
      cudaEventDestroy(io->filled[s]);
      cudaEventDestroy(io->drained[s]);
    }
  }
  if (io->d_capoff)
    cudaFree(io->d_capoff);
  if (io->h_capoff)
    cudaFreeHost(io->h_capoff);
  if (io->ev_misc)
    cudaEventDestroy(io->ev_misc);
  if (io->h2d)
    cudaStreamDestroy(io->h2d);
  if (io->d2h)
    cudaStreamDestroy(io->d2h);
  delete io;
  a->hostio = nullptr;
}

namespace
{

int hostio_get(picnix_arena* a, int64_t min_slab_elems, HostIO** out)
{
  if (a->hostio == nullptr)
    a->hostio = new HostIO();
  HostIO* io = a->hostio;
  if (!io->ready) {
    PICNIX_CUDA(a, cudaStreamCreateWithFlags(&io->h2d, cudaStreamNonBlocking));
    PICNIX_CUDA(a, cudaStreamCreateWithFlags(&io->d2h, cudaStreamNonBlocking));
    for (int s = 0; s < HostIO::NSLOT; s++) {
      PICNIX_CUDA(a, cudaEventCreateWithFlags(&io->filled[s], cudaEventDisableTiming));
      PICNIX_CUDA(a, cudaEventCreateWithFlags(&io->drained[s], cudaEventDisableTiming));
    }
    PICNIX_CUDA(a, cudaEventCreateWithFlags(&io->ev_misc, cudaEventDisableTiming));
    io->ready = true;
  }
  // slabs of 256 MB keep the copy engines busy with few, large transfers; a segment never straddles
  // two slabs, so a slab is at least one segment long
  const int64_t want = std::max<int64_t>(min_slab_elems, (int64_t)32 << 20);
  if (io->slab_elems < want) {
    for (int s = 0; s < HostIO::NSLOT; s++) {
      if (io->slab[s])
        cudaFree(io->slab[s]);
      io->slab[s] = nullptr;
    }
    io->slab_elems = 0;
    for (int s = 0; s < HostIO::NSLOT; s++)
      PICNIX_CUDA(a, cudaMalloc((void**)&io->slab[s], want * sizeof(double)));
    io->slab_elems = want;
  }
  *out = io;
  return PICNIX_OK;
}

// page-lock a caller's buffer unless it already is (cudaMallocHost / cudaHostRegister / torch pinned)
void pin_if_needed(HostIO* io, const void* ptr, size_t bytes)
{
  if (ptr == nullptr || bytes == 0)
    return;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, ptr) == cudaSuccess && attr.type != cudaMemoryTypeUnregistered)
    return;
  cudaGetLastError();
  if (cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterDefault) == cudaSuccess)
    io->registered.push_back(const_cast<void*>(ptr));
  else
    cudaGetLastError(); // stays pageable: the copies still work, only slower
}

// one piece of a batch: `elems` doubles at host address `host`, staged at slab offset `slab_off`
struct Piece {
  double* host;
  int64_t elems;
  int64_t slab_off;
  int     id; // segment (particles) or first chunk (ff)
  int     n;  // particles / chunks
  int     group = 0; // > 0: `id` is the first of `group` consecutive segments copied as one host span
  int     maxn  = 0; // largest particle count in the group
};

// Arenas with very many segments (1-D runs) move their particles in groups of consecutive segments:
// one copy of the host span (capacity gaps included) and one transposition launch per group instead of
// one of each per segment (two-stream, 98 304 segments: 1.6 s -> tens of ms per step).
constexpr int     GROUP_MIN_SEGMENTS = 2048;
constexpr int     GROUP_MAX_SEGMENTS = 32768;      // blockIdx.y
constexpr int64_t GROUP_MAX_ELEMS    = 8 << 20;    // 64 MB per copy keeps the three streams busy

// particle pieces of the arena: per segment, or per group of segments
static void particle_pieces(picnix_arena* a, HostIO* io, double* xu, const int32_t* np, const int32_t* np_cap,
                            std::vector<Piece>& pieces)
{
  if (a->nseg < GROUP_MIN_SEGMENTS) {
    int64_t poff = 0;
    for (int s = 0; s < a->nseg; s++) {
      pieces.push_back({xu + poff * NC, (int64_t)np[s] * NC, 0, s, np[s]});
      poff += np_cap[s];
    }
    return;
  }
  const int64_t limit = std::min<int64_t>(io->slab_elems, GROUP_MAX_ELEMS);
  int           s0    = 0;
  while (s0 < a->nseg) {
    int     s1 = s0, maxn = np[s0];
    int64_t span = (int64_t)np[s0] * NC;
    while (s1 + 1 < a->nseg && s1 + 1 - s0 < GROUP_MAX_SEGMENTS) {
      const int64_t next = (io->h_capoff[s1 + 1] - io->h_capoff[s0] + np[s1 + 1]) * NC;
      if (next > limit)
        break;
      s1++;
      span = next;
      maxn = std::max(maxn, np[s1]);
    }
    Piece p{xu + io->h_capoff[s0] * NC, span, 0, s0, 0};
    p.group = s1 - s0 + 1;
    p.maxn  = maxn;
    pieces.push_back(p);
    s0 = s1 + 1;
  }
}

// split pieces into batches that fit a slab
std::vector<std::vector<Piece>> make_batches(std::vector<Piece>& pieces, int64_t slab_elems)
{
  std::vector<std::vector<Piece>> batches;
  std::vector<Piece>              cur;
  int64_t                         fill = 0;
  for (Piece& p : pieces) {
    if (p.elems == 0)
      continue;
    if (fill + p.elems > slab_elems && !cur.empty()) {
      batches.push_back(cur);
      cur.clear();
      fill = 0;
    }
    p.slab_off = fill;
    fill += p.elems;
    cur.push_back(p);
  }
  if (!cur.empty())
    batches.push_back(cur);
  return batches;
}

constexpr int TTHREADS = 256;

} // namespace

// common prologue: slabs, streams, page-locking of the caller's arrays
static int hostio_prepare(picnix_arena* a, double* uf, double* uj, double* ff, double* xu,
                          const int32_t* np_cap, HostIO** out)
{
  const Geom&   g     = a->g;
  const int64_t ncell = g.Ng;
  int           status;
  if (!a->particles_allocated) {
    if ((status = picnix_cuda_set_particle_capacity(a, np_cap)) != PICNIX_OK)
      return status;
  }
  int64_t cap_total = 0, max_seg = 0;
  for (int s = 0; s < a->nseg; s++) {
    cap_total += np_cap[s];
    max_seg = std::max<int64_t>(max_seg, (int64_t)a->seg_cap[s] * NC);
  }
  max_seg = std::max<int64_t>(max_seg, ncell * 18);
  HostIO* io = nullptr;
  if ((status = hostio_get(a, max_seg, &io)) != PICNIX_OK)
    return status;
  if (a->nseg >= GROUP_MIN_SEGMENTS) {
    if (io->capoff_len < a->nseg + 1) {
      if (io->d_capoff)
        cudaFree(io->d_capoff);
      if (io->h_capoff)
        cudaFreeHost(io->h_capoff);
      io->d_capoff = nullptr;
      io->h_capoff = nullptr;
      PICNIX_CUDA(a, cudaMalloc((void**)&io->d_capoff, (a->nseg + 1) * sizeof(int64_t)));
      PICNIX_CUDA(a, cudaMallocHost((void**)&io->h_capoff, (a->nseg + 1) * sizeof(int64_t)));
      io->capoff_len = a->nseg + 1;
    }
    io->h_capoff[0] = 0;
    for (int s = 0; s < a->nseg; s++)
      io->h_capoff[s + 1] = io->h_capoff[s] + np_cap[s];
    // on the compute stream, which every transposition launch is ordered behind
    PICNIX_CUDA(a, cudaMemcpyAsync(io->d_capoff, io->h_capoff, (a->nseg + 1) * sizeof(int64_t),
                                   cudaMemcpyHostToDevice, a->stream));
  }
  pin_if_needed(io, uf, (size_t)g.nchunk * ncell * 6 * sizeof(double));
  pin_if_needed(io, uj, (size_t)g.nchunk * ncell * 4 * sizeof(double));
  pin_if_needed(io, ff, (size_t)g.nchunk * ncell * 18 * sizeof(double));
  pin_if_needed(io, xu, (size_t)cap_total * NC * sizeof(double));
  *out = io;
  return PICNIX_OK;
}

// host arrays -> device state; uploaded particles are counted and sorted so that pindex is valid
int upload_state_pipelined(picnix_arena* a, double* uf, double* uj, double* ff, double* xu,
                           const int32_t* np_in, const int32_t* np_cap, bool with_uj)
{
  const Geom&   g     = a->g;
  const int64_t ncell = g.Ng;
  int           status;
  HostIO*       io = nullptr;
  if ((status = hostio_prepare(a, uf, uj, ff, xu, np_cap, &io)) != PICNIX_OK)
    return status;
  for (int s = 0; s < a->nseg; s++)
    if (np_in[s] < 0 || np_in[s] > a->seg_cap[s])
      return fail(a, PICNIX_ERR_OVERFLOW, "upload_state: np_in exceeds segment capacity");

  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  for (int s = 0; s < HostIO::NSLOT; s++)
    io->used[s] = false;

  PICNIX_CUDA(a, cudaMemcpyAsync(a->d.uf, uf, (size_t)g.nchunk * ncell * 6 * sizeof(double),
                                 cudaMemcpyHostToDevice, io->h2d));
  if (with_uj)
    PICNIX_CUDA(a, cudaMemcpyAsync(a->d.uj, uj, (size_t)g.nchunk * ncell * 4 * sizeof(double),
                                   cudaMemcpyHostToDevice, io->h2d));
  PICNIX_CUDA(a, cudaMemcpyAsync(a->d.np, np_in, a->nseg * sizeof(int), cudaMemcpyHostToDevice,
                                 io->h2d));
  PICNIX_CUDA(a, cudaMemsetAsync(a->d.ntail, 0, a->nseg * sizeof(int), io->h2d));

  std::vector<Piece> pieces;
  // ff: whole chunks, as many as fit a slab
  {
    const int per = (int)std::max<int64_t>(1, io->slab_elems / (ncell * 18));
    for (int c = 0; c < g.nchunk; c += per) {
      int n = std::min(per, g.nchunk - c);
      pieces.push_back({ff + (int64_t)c * ncell * 18, (int64_t)n * ncell * 18, 0, -1 - c, n});
    }
  }
  particle_pieces(a, io, xu, np_in, np_cap, pieces);
  auto batches = make_batches(pieces, io->slab_elems);

  int slot = 0;
  for (auto& batch : batches) {
    if (io->used[slot])
      PICNIX_CUDA(a, cudaStreamWaitEvent(io->h2d, io->drained[slot], 0));
    for (const Piece& p : batch)
      PICNIX_CUDA(a, cudaMemcpyAsync(io->slab[slot] + p.slab_off, p.host, p.elems * sizeof(double),
                                     cudaMemcpyHostToDevice, io->h2d));
    PICNIX_CUDA(a, cudaEventRecord(io->filled[slot], io->h2d));
    PICNIX_CUDA(a, cudaStreamWaitEvent(a->stream, io->filled[slot], 0));
    for (const Piece& p : batch) {
      if (p.id < 0) {
        const int     c0 = -1 - p.id;
        const int64_t nc = (int64_t)p.n * ncell;
        ff_compact_kernel<<<(unsigned)((nc * 9 + TTHREADS - 1) / TTHREADS), TTHREADS, 0, a->stream>>>(
            io->slab[slot] + p.slab_off, a->d.ff + (int64_t)c0 * ncell * 9, nc);
      } else if (p.group > 0) {
        if (p.maxn > 0) {
          const dim3 grid((unsigned)(((int64_t)p.maxn * NC + TTHREADS - 1) / TTHREADS), (unsigned)p.group);
          aos_to_soa_group_kernel<<<grid, TTHREADS, 0, a->stream>>>(io->slab[slot] + p.slab_off, a->d.xu,
                                                                    io->d_capoff, a->d.seg_off, a->d.np, p.id,
                                                                    a->d.pcap);
        }
      } else {
        aos_to_soa_kernel<<<(unsigned)((p.elems + TTHREADS - 1) / TTHREADS), TTHREADS, 0, a->stream>>>(
            io->slab[slot] + p.slab_off, a->d.xu, a->seg_off[p.id], a->d.pcap, p.n);
      }
      a->kernel_launches++;
    }
    PICNIX_CUDA(a, cudaEventRecord(io->drained[slot], a->stream));
    io->used[slot] = true;
    slot           = (slot + 1) % HostIO::NSLOT;
  }
  // the direct copies (uf, np) must have landed before the first kernel that follows
  PICNIX_CUDA(a, cudaEventRecord(io->ev_misc, io->h2d));
  PICNIX_CUDA(a, cudaStreamWaitEvent(a->stream, io->ev_misc, 0));
  PICNIX_CUDA(a, cudaGetLastError());

  a->pindex_valid     = false;
  a->leave_list_valid = false;
  a->perm_pending     = false; // the arrays were overwritten: nothing left to reorder
  if ((status = launch_count(a, 0, -1)) != PICNIX_OK)
    return status;
  return launch_sort(a, 0, -1);
}

// device state -> host arrays (np_out receives the particle counts)
int download_state_pipelined(picnix_arena* a, double* uf, double* uj, double* ff, double* xu,
                             const int32_t* np_cap, int32_t* np_out)
{
  const Geom&   g     = a->g;
  const int64_t ncell = g.Ng;
  int           status;
  HostIO*       io = nullptr;
  if ((status = hostio_prepare(a, uf, uj, ff, xu, np_cap, &io)) != PICNIX_OK)
    return status;
  if ((status = materialize_sort(a)) != PICNIX_OK)
    return status;
  if ((status = picnix_cuda_get_np(a, np_out)) != PICNIX_OK) // synchronises the compute stream
    return status;
  for (int s = 0; s < a->nseg; s++)
    if (np_out[s] > np_cap[s])
      return fail(a, PICNIX_ERR_OVERFLOW, "host particle buffer too small for the new count");
  PICNIX_CUDA(a, cudaStreamSynchronize(io->h2d));
  for (int s = 0; s < HostIO::NSLOT; s++)
    io->used[s] = false;

  PICNIX_CUDA(a, cudaMemcpyAsync(uf, a->d.uf, (size_t)g.nchunk * ncell * 6 * sizeof(double),
                                 cudaMemcpyDeviceToHost, io->d2h));
  PICNIX_CUDA(a, cudaMemcpyAsync(uj, a->d.uj, (size_t)g.nchunk * ncell * 4 * sizeof(double),
                                 cudaMemcpyDeviceToHost, io->d2h));

  std::vector<Piece> pieces;
  particle_pieces(a, io, xu, np_out, np_cap, pieces);
  {
    const int per = (int)std::max<int64_t>(1, io->slab_elems / (ncell * 18));
    for (int c = 0; c < g.nchunk; c += per) {
      int n = std::min(per, g.nchunk - c);
      pieces.push_back({ff + (int64_t)c * ncell * 18, (int64_t)n * ncell * 18, 0, -1 - c, n});
    }
  }
  auto batches = make_batches(pieces, io->slab_elems);

  int slot = 0;
  for (auto& batch : batches) {
    if (io->used[slot])
      PICNIX_CUDA(a, cudaStreamWaitEvent(a->stream, io->drained[slot], 0));
    for (const Piece& p : batch) {
      if (p.id < 0) {
        const int     c0 = -1 - p.id;
        const int64_t nc = (int64_t)p.n * ncell;
        ff_expand_kernel<<<(unsigned)((nc * 18 + TTHREADS - 1) / TTHREADS), TTHREADS, 0, a->stream>>>(
            a->d.ff + (int64_t)c0 * ncell * 9, io->slab[slot] + p.slab_off, nc);
      } else if (p.group > 0) {
        if (p.maxn > 0) {
          const dim3 grid((unsigned)(((int64_t)p.maxn * NC + TTHREADS - 1) / TTHREADS), (unsigned)p.group);
          soa_to_aos_group_kernel<<<grid, TTHREADS, 0, a->stream>>>(a->d.xu, io->slab[slot] + p.slab_off,
                                                                    io->d_capoff, a->d.seg_off, a->d.np, p.id,
                                                                    a->d.pcap);
        }
      } else {
        soa_to_aos_kernel<<<(unsigned)((p.elems + TTHREADS - 1) / TTHREADS), TTHREADS, 0, a->stream>>>(
            a->d.xu, io->slab[slot] + p.slab_off, a->seg_off[p.id], a->d.pcap, p.n);
      }
      a->kernel_launches++;
    }
    PICNIX_CUDA(a, cudaEventRecord(io->filled[slot], a->stream));
    PICNIX_CUDA(a, cudaStreamWaitEvent(io->d2h, io->filled[slot], 0));
    for (const Piece& p : batch)
      PICNIX_CUDA(a, cudaMemcpyAsync(p.host, io->slab[slot] + p.slab_off, p.elems * sizeof(double),
                                     cudaMemcpyDeviceToHost, io->d2h));
    PICNIX_CUDA(a, cudaEventRecord(io->drained[slot], io->d2h));
    io->used[slot] = true;
    slot           = (slot + 1) % HostIO::NSLOT;
  }
  PICNIX_CUDA(a, cudaGetLastError());
  PICNIX_CUDA(a, cudaStreamSynchronize(io->d2h));
  return picnix_cuda_synchronize(a);
}

int step_host_pipelined(picnix_arena* a, double delt, int nstep, double* uf, double* uj, double* ff,
                        double* xu, const int32_t* np_in, const int32_t* np_cap, int32_t* np_out)
{
  int status = upload_state_pipelined(a, uf, uj, ff, xu, np_in, np_cap, nstep < 1);
  if (status != PICNIX_OK)
    return status;
  if ((status = picnix_cuda_step(a, delt, nstep)) != PICNIX_OK)
    return status;
  return download_state_pipelined(a, uf, uj, ff, xu, np_cap, np_out);
}

} // namespace picnix

extern "C" {

int picnix_cuda_upload_state(picnix_arena_t* a, double* uf, double* uj, double* ff, double* xu,
                             const int32_t* np_in, const int32_t* np_cap)
{
  if (a == nullptr || uf == nullptr || uj == nullptr || ff == nullptr || xu == nullptr ||
      np_in == nullptr || np_cap == nullptr)
    return PICNIX_ERR_INVALID;
  return picnix::upload_state_pipelined(a, uf, uj, ff, xu, np_in, np_cap, true);
}

int picnix_cuda_download_state(picnix_arena_t* a, double* uf, double* uj, double* ff, double* xu,
                               const int32_t* np_cap, int32_t* np_out)
{
  if (a == nullptr || uf == nullptr || uj == nullptr || ff == nullptr || xu == nullptr ||
      np_cap == nullptr || np_out == nullptr)
    return PICNIX_ERR_INVALID;
  return picnix::download_state_pipelined(a, uf, uj, ff, xu, np_cap, np_out);
}

int picnix_cuda_host_alloc(void** ptr, int64_t bytes)
{
  if (ptr == nullptr || bytes < 0)
    return PICNIX_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    cudaGetLastError();
    return PICNIX_ERR_NODEVICE;
  }
  return cudaMallocHost(ptr, (size_t)std::max<int64_t>(bytes, 1)) == cudaSuccess ? PICNIX_OK
                                                                                   : PICNIX_ERR_CUDA;
}

int picnix_cuda_host_free(void* ptr)
{
  if (ptr == nullptr)
    return PICNIX_OK;
  return cudaFreeHost(ptr) == cudaSuccess ? PICNIX_OK : PICNIX_ERR_CUDA;
}

} // extern "C"
