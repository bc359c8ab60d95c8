// -*- C++ -*-
// K4: counting sort that keeps particles cell-ordered, batched over all (chunk, species) segments.
//
// Reference: XtensorParticle::sort (nix/xtensor_particle.hpp:260-321).  The reference stripes its
// histogram over NIX_SIMD_WIDTH lanes and scatters in the order (cell, ip % 8, ip); only
// `pindex`, `Np` and the SET of particles in each cell are implementation independent
// (SURVEY Appendix A), and those are what this kernel reproduces bit-exactly:
//   pindex[k] = number of particles with key < k   (k = 0..Ng),   Np = pindex[Ng]
// Particles whose key is Ng (outside the chunk) are dropped.
//
// Three steps, all on the arena's stream, no host round trip:
//   1. exclusive scan of the per-segment histogram pcount -> pindex, and pcount <- pindex (cursor)
//   2. scatter xu -> xv with one atomic cursor bump per particle (order inside a cell is free)
//   3. Np <- pindex[Ng], tail counter <- 0; the host swaps the xu/xv pointers
//
// Lazy variant (option "lazy_sort", on by default, used when the tiled push kernel will consume the
// result): step 2 writes only the permutation `perm[sorted slot] = current slot` (8 B per particle
// instead of 116 B).  The next fused push reads its particles through `perm` and writes them to
// the other buffer in sorted order, so the physical reordering rides on traffic the push has
// anyway.  Any other consumer of the particle arrays (downloads, the generic kernels, moments,
// chunk moves) first calls materialize_sort(), which performs the gather the scatter would have
// done.  pindex, Np and the per-cell particle sets are identical in both variants.
#include "arena.hpp"

#include <cstdlib>

namespace picnix
{

namespace
{

constexpr int SCAN_THREADS    = 256;
constexpr int SCATTER_THREADS = 256;

__device__ __forceinline__ int warp_inclusive_scan(int v)
{
#pragma unroll
  for (int ofs = 1; ofs < 32; ofs <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, v, ofs);
    if ((threadIdx.x & 31) >= ofs)
      v += n;
  }
  return v;
}

// one block per segment; bins are consumed in tiles of SCAN_THREADS with a running carry
__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(Geom g, DevPtrs d, int seg0)
{
  __shared__ int warp_sum[SCAN_THREADS / 32];
  __shared__ int carry;

  const int     seg  = seg0 + blockIdx.x;
  const int     nbin = g.Ng + 1;
  int*          cnt  = d.pcount + (int64_t)seg * nbin;
  int*          pix  = d.pindex + (int64_t)seg * nbin;
  const int     lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (threadIdx.x == 0)
    carry = 0;
  __syncthreads();

  for (int base = 0; base < nbin; base += SCAN_THREADS) {
    int k = base + threadIdx.x;
    int c = k < nbin ? cnt[k] : 0;
    int s = warp_inclusive_scan(c);
    if (lane == 31)
      warp_sum[warp] = s;
    __syncthreads();
    if (warp == 0) {
      int w = lane < SCAN_THREADS / 32 ? warp_sum[lane] : 0;
      w     = warp_inclusive_scan(w);
      if (lane < SCAN_THREADS / 32)
        warp_sum[lane] = w;
    }
    __syncthreads();
    int prefix = carry + (warp > 0 ? warp_sum[warp - 1] : 0) + s - c; // exclusive
    if (k < nbin) {
      pix[k] = prefix;
      cnt[k] = prefix; // scatter cursor
    }
    __syncthreads();
    if (threadIdx.x == SCAN_THREADS - 1)
      carry = prefix + c;
    __syncthreads();
  }
}

// Few bins per segment and many segments (1-D boxes: 301-501 bins, 10^4-10^5 segments): one WARP per segment, eight
// segments per block.  All bins of the segment are loaded before the first scan (NITER independent loads in
// flight per lane), then scanned 32 at a time with a running carry -- no block barrier, and one memory
// latency per segment instead of one per 256 bins.
template <int NITER>
__global__ void __launch_bounds__(SCAN_THREADS) scan_warp_kernel(Geom g, DevPtrs d, int seg0, int nseg)
{
  const int lane = threadIdx.x & 31;
  const int wseg = blockIdx.x * (SCAN_THREADS / 32) + (threadIdx.x >> 5);
  if (wseg >= nseg)
    return;
  const int      seg  = seg0 + wseg;
  const int      nbin = g.Ng + 1;
  int* __restrict__ cnt = d.pcount + (int64_t)seg * nbin;
  int* __restrict__ pix = d.pindex + (int64_t)seg * nbin;
  int            c[NITER];
#pragma unroll
  for (int i = 0; i < NITER; i++) {
    const int k = i * 32 + lane;
    c[i]        = k < nbin ? cnt[k] : 0;
  }
  int carry = 0;
#pragma unroll
  for (int i = 0; i < NITER; i++) {
    const int k = i * 32 + lane;
    const int s = warp_inclusive_scan(c[i]);
    const int e = carry + s - c[i]; // exclusive
    if (k < nbin) {
      pix[k] = e;
      cnt[k] = e; // scatter cursor
    }
    carry += __shfl_sync(0xffffffffu, s, 31);
  }
}

// One thread per particle.  Cell-ordered input means the lanes of a warp hold a handful of distinct
// keys, so the cursor is bumped once per distinct key and warp (match.any + one leader atomic)
// instead of once per particle: ~30x fewer same-address L2 atomics at 32 particles per cell.
__global__ void __launch_bounds__(SCATTER_THREADS)
scatter_kernel(Geom g, DevPtrs d, int seg0, int blocks_per_seg)
{
  const int lseg = blockIdx.x / blocks_per_seg;
  const int b    = blockIdx.x - lseg * blocks_per_seg;
  const int seg  = seg0 + lseg;
  const int ip   = b * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  // own particles plus the migrants appended behind them during this step
  const int n    = min(d.np[seg] + d.ntail[seg], d.seg_cap[seg]);
  if (b * (int)blockDim.x + (int)(threadIdx.x & ~31u) >= n)
    return; // whole warp beyond the segment

  const int64_t off = d.seg_off[seg];
  // particles that left the chunk (key == Ng) are discarded (nix/xtensor_particle.hpp:319-320)
  const int  key  = ip < n ? d.gindex[off + ip] : g.Ng;
  const bool keep = key < g.Ng;

  const unsigned peers  = __match_any_sync(0xffffffffu, key);
  const int      leader = __ffs(peers) - 1;
  int            base   = 0;
  if (keep && lane == leader)
    base = atomicAdd(d.pcount + (int64_t)seg * (g.Ng + 1) + key, __popc(peers));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (!keep)
    return;
  const int dst = base + __popc(peers & ((1u << lane) - 1u));
#pragma unroll
  for (int k = 0; k < NC; k++)
    d.xv[k * d.pcap + off + dst] = d.xu[k * d.pcap + off + ip];
}

// lazy sort: same cursor logic, but only the permutation is written.  The kernel is a chain of dependent
// round trips (key load -> match -> L2 atomic -> store), so every warp keeps INDEX_ILP independent chains
// in flight: it owns INDEX_ILP * 32 consecutive particles, lane-contiguous per chain.
constexpr int INDEX_ILP = 4;

__global__ void __launch_bounds__(SCATTER_THREADS)
scatter_index_kernel(Geom g, DevPtrs d, int seg0, int blocks_per_seg)
{
  const int lseg = blockIdx.x / blocks_per_seg;
  const int b    = blockIdx.x - lseg * blocks_per_seg;
  const int seg  = seg0 + lseg;
  const int lane = threadIdx.x & 31;
  const int n    = min(d.np[seg] + d.ntail[seg], d.seg_cap[seg]);
  // first particle of this warp
  const int w0   = (b * (SCATTER_THREADS / 32) + (threadIdx.x >> 5)) * (32 * INDEX_ILP);
  if (w0 >= n)
    return;

  const int64_t off = d.seg_off[seg];
  int*          cur = d.pcount + (int64_t)seg * (g.Ng + 1);
  int           key[INDEX_ILP], base[INDEX_ILP];
  unsigned      peers[INDEX_ILP];
#pragma unroll
  for (int j = 0; j < INDEX_ILP; j++) {
    const int ip = w0 + j * 32 + lane;
    key[j]       = ip < n ? d.gindex[off + ip] : g.Ng;
  }
#pragma unroll
  for (int j = 0; j < INDEX_ILP; j++) {
    peers[j] = __match_any_sync(0xffffffffu, key[j]);
    base[j]  = 0;
    if (key[j] < g.Ng && lane == __ffs(peers[j]) - 1)
      base[j] = atomicAdd(cur + key[j], __popc(peers[j]));
  }
#pragma unroll
  for (int j = 0; j < INDEX_ILP; j++) {
    base[j] = __shfl_sync(0xffffffffu, base[j], __ffs(peers[j]) - 1);
    if (key[j] < g.Ng)
      d.perm[off + base[j] + __popc(peers[j] & ((1u << lane) - 1u))] = w0 + j * 32 + lane;
  }
}

// materialise a pending lazy sort: xv[sorted slot] <- xu[perm[sorted slot]]
__global__ void __launch_bounds__(SCATTER_THREADS)
gather_kernel(Geom g, DevPtrs d, int seg0, int blocks_per_seg)
{
  const int lseg = blockIdx.x / blocks_per_seg;
  const int b    = blockIdx.x - lseg * blocks_per_seg;
  const int seg  = seg0 + lseg;
  const int j    = b * blockDim.x + threadIdx.x;
  if (j >= d.np[seg])
    return;
  const int64_t off = d.seg_off[seg];
  const int     src = d.perm[off + j];
#pragma unroll
  for (int k = 0; k < NC; k++)
    d.xv[k * d.pcap + off + j] = d.xu[k * d.pcap + off + src];
}

__global__ void finish_kernel(Geom g, DevPtrs d, int seg0, int nseg)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nseg)
    return;
  int seg   = seg0 + t;
  d.np[seg] = d.pindex[(int64_t)seg * (g.Ng + 1) + g.Ng];
  // ntail (the arrivals of this step) is read by the growth statistics and cleared there
}

} // namespace

// XtensorParticle::sort for segments of chunks [c0, c0+cn); the histogram must be current
int launch_sort(picnix_arena* a, int c0, int cn)
{
  resolve_range(a, c0, cn);
  {
    // a sort on top of a pending index sort (e.g. count + sort after a step): order physically first
    int status = materialize_sort(a);
    if (status != PICNIX_OK)
      return status;
  }
  if (!a->particles_allocated)
    return fail(a, PICNIX_ERR_INVALID, "no particles allocated");
  if (cn == 0)
    return PICNIX_OK;
  if (c0 != 0 || cn != a->g.nchunk)
    return fail(a, PICNIX_ERR_INVALID,
                "sort_particle acts on all chunks of the arena (xu/xv are swapped arena-wide)");

  const Geom& g    = a->g;
  const int   seg0 = c0 * g.Ns;
  const int   nseg = cn * g.Ns;

  {
    const int nbin   = g.Ng + 1;
    const int wblock = (nseg + SCAN_THREADS / 32 - 1) / (SCAN_THREADS / 32);
    // measured on one box (PICNIX_SCAN_WARP=0 restores the block scan): two-stream (301 bins, 98 304 segments)
    // 1.32 -> 1.21 ms per step, shock (501 bins) 1.157 -> 1.125 ms; 2-D boxes keep the block scan
    static const bool warp_scan = !(std::getenv("PICNIX_SCAN_WARP") && std::atoi(std::getenv("PICNIX_SCAN_WARP")) == 0);
    if (warp_scan && g.dimension == 1 && nbin <= 10 * 32)
      scan_warp_kernel<10><<<wblock, SCAN_THREADS, 0, a->stream>>>(g, a->d, seg0, nseg);
    else if (warp_scan && g.dimension == 1 && nbin <= 16 * 32)
      scan_warp_kernel<16><<<wblock, SCAN_THREADS, 0, a->stream>>>(g, a->d, seg0, nseg);
    else
      scan_kernel<<<nseg, SCAN_THREADS, 0, a->stream>>>(g, a->d, seg0);
  }
  a->kernel_launches++;

  int maxcap = 0;
  for (int s = seg0; s < seg0 + nseg; s++)
    maxcap = std::max(maxcap, a->seg_cap[s]);
  int bps = (maxcap + SCATTER_THREADS - 1) / SCATTER_THREADS;
  // the tiled push kernel reads through the permutation: no need to move the particles now
  const bool lazy = a->lazy_sort && !a->force_generic && row_geometry_applies(a);
  if (bps > 0) {
    if (lazy) {
      const int bpi = (maxcap + SCATTER_THREADS * INDEX_ILP - 1) / (SCATTER_THREADS * INDEX_ILP);
      scatter_index_kernel<<<bpi * nseg, SCATTER_THREADS, 0, a->stream>>>(g, a->d, seg0, bpi);
    }
    else
      scatter_kernel<<<bps * nseg, SCATTER_THREADS, 0, a->stream>>>(g, a->d, seg0, bps);
    a->kernel_launches++;
  }
  finish_kernel<<<(nseg + 127) / 128, 128, 0, a->stream>>>(g, a->d, seg0, nseg);
  a->kernel_launches++;
  PICNIX_CUDA(a, cudaGetLastError());
  {
    // smallest free space / largest number of arrivals / spill count: the next step's growth decision
    int status = record_segment_stats(a);
    if (status != PICNIX_OK)
      return status;
  }

  if (lazy) {
    a->perm_pending = true;
  } else {
    // XtensorParticle::swap (nix/xtensor_particle.hpp:120-123)
    std::swap(a->d.xu, a->d.xv);
  }
  a->pindex_valid     = true;
  a->leave_list_valid = false; // slots changed
  return PICNIX_OK;
}

int materialize_sort(picnix_arena* a)
{
  if (!a->perm_pending)
    return PICNIX_OK;
  const Geom& g      = a->g;
  int         maxcap = 0;
  for (int s = 0; s < a->nseg; s++)
    maxcap = std::max(maxcap, a->seg_cap[s]);
  const int bps = (maxcap + SCATTER_THREADS - 1) / SCATTER_THREADS;
  if (bps > 0) {
    gather_kernel<<<bps * a->nseg, SCATTER_THREADS, 0, a->stream>>>(g, a->d, 0, bps);
    a->kernel_launches++;
  }
  PICNIX_CUDA(a, cudaGetLastError());
  std::swap(a->d.xu, a->d.xv);
  a->perm_pending = false;
  return PICNIX_OK;
}

} // namespace picnix
