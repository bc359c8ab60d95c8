// -*- C++ -*-
// Growing particle storage: the device side of XtensorParticle::resize (nix/xtensor_particle.hpp:70-115)
// as XtensorHaloParticle3D::pre_unpack uses it (nix/xtensor_halo3d.hpp:406-418: every step the arrays of a
// chunk are resized to hold the particles about to arrive, BEFORE they are unpacked).
//
// All (chunk, species) segments of a rank live in shared SoA arrays with fixed offsets, so a segment
// cannot grow in place.  Instead:
//   * a migrant that finds its destination segment full is not dropped: append_particle() puts it on
//     the spill list (migrate.cuh);
//   * resolve_growth() runs between the unpack of the particle exchange and the sort.  When the host
//     has to look (see below) it reads the segment populations and the spill list, re-lays out the
//     particle arrays with larger segments where needed (new capacity = population x (1 + buffer_ratio),
//     rounded like Particle::round_up_alloc) and appends the spilled migrants; the sort then sees exactly
//     the particles the reference's sort would see.  Nothing is lost, nothing aborts.
//   * looking costs a host synchronisation in the middle of the step, so it is done only when the
//     statistics of the PREVIOUS step (smallest free space over all segments, largest number of arrivals
//     in a segment, spill count; they return from the device one step late through pinned memory) say a
//     segment might fill up: free < 2 x arrivals + 16, or something was spilled.  A
//     Courant-limited flow cannot go from "plenty of room" to "full" faster than that.  If it happens
//     anyway the spilled migrants are appended one step late and counted (picnix_cuda_get_growth_stats).
//   * a full message to a peer (lagged-count protocol) spills the same way; those records leave with the
//     next exchange and are counted as late as well.
//
// Peak memory of a re-layout equals the steady state: the scratch buffer xv is released first, the new
// xu allocated, the segments copied, the old xu released, the new xv allocated.
#include "migrate.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace picnix
{

namespace
{

constexpr int GROW_THREADS = 256;

__global__ void __launch_bounds__(GROW_THREADS)
copy_segments_kernel(DevPtrs d, const int64_t* __restrict__ new_off, double* __restrict__ new_xu,
                     int* __restrict__ new_gindex, int64_t new_pcap, int blocks_per_seg)
{
  const int seg = blockIdx.x / blocks_per_seg;
  const int n   = d.np[seg] + d.ntail[seg];
  const int64_t src0 = d.seg_off[seg], dst0 = new_off[seg];
  for (int j = (blockIdx.x - seg * blocks_per_seg) * blockDim.x + threadIdx.x; j < n;
       j += blocks_per_seg * blockDim.x) {
#pragma unroll
    for (int k = 0; k < NC; k++)
      new_xu[k * new_pcap + dst0 + j] = d.xu[k * d.pcap + src0 + j];
    new_gindex[dst0 + j] = d.gindex[src0 + j];
  }
}

// spilled migrants with a local destination -> behind the active particles of their (grown) segment;
// records for remote neighbours are compacted to the front of `keep`
__global__ void __launch_bounds__(GROW_THREADS)
drain_spill_kernel(Geom g, DevPtrs d, int n, double* __restrict__ keep, int* __restrict__ nkeep)
{
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const double* in = d.spill_rec + (int64_t)r * 8;
    double        p[NC];
#pragma unroll
    for (int k = 0; k < NC; k++)
      p[k] = in[k];
    double     tagbits = in[7];
    const int2 tag     = *reinterpret_cast<int2*>(&tagbits);
    if (tag.x >= 0) {
      const int chunk = tag.x / g.Ns;
      append_particle(g, d, chunk, tag.x - chunk * g.Ns, p);
    } else {
      const int k  = atomicAdd(nkeep, 1);
      double*   out = keep + (int64_t)k * 8;
#pragma unroll
      for (int c = 0; c < 8; c++)
        out[c] = in[c];
    }
  }
}

__global__ void segment_stat_kernel(DevPtrs d, int nseg)
{
  // after the sort: np is final, ntail still holds the arrivals of this step (reset right after)
  int minfree = 0x7fffffff, maxtail = 0;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nseg; s += gridDim.x * blockDim.x) {
    minfree = min(minfree, d.seg_cap[s] - d.np[s]);
    maxtail = max(maxtail, d.ntail[s]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    minfree = min(minfree, __shfl_xor_sync(0xffffffffu, minfree, o));
    maxtail = max(maxtail, __shfl_xor_sync(0xffffffffu, maxtail, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(d.seg_stat + 0, minfree);
    atomicMax(d.seg_stat + 1, maxtail);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    d.seg_stat[2] = *d.spill_count;
    // fatal device-side flags ride along (bit 0: spill list full, bit 2: foreign particle received,
    // bit 3: far-mover list full), so a free-running host learns of them one step late without a
    // synchronisation of its own
    d.seg_stat[3] = (d.errflag[0] != 0 ? 1 : 0) | (d.errflag[2] != 0 ? 4 : 0) | (d.errflag[3] != 0 ? 8 : 0);
  }
}

__global__ void reset_tail_kernel(DevPtrs d, int nseg)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nseg)
    d.ntail[s] = 0;
}

template <typename T>
int alloc(picnix_arena* a, T** ptr, size_t count)
{
  PICNIX_CUDA(a, cudaMalloc((void**)ptr, std::max<size_t>(count, 1) * sizeof(T)));
  return PICNIX_OK;
}

// new layout with capacities `cap` (already rounded); the first np + ntail slots of every segment survive
int relayout(picnix_arena* a, const std::vector<int32_t>& cap)
{
  std::vector<int64_t> off(a->nseg);
  int64_t              total = 0;
  for (int s = 0; s < a->nseg; s++) {
    off[s] = total;
    total += cap[s];
  }
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  // scratch first: xv and the lazy-sort permutation hold nothing between the exchange and the sort
  cudaFree(a->d.xv);
  cudaFree(a->d.perm);
  a->d.xv   = nullptr;
  a->d.perm = nullptr;

  double*  new_xu = nullptr;
  int*     new_gi = nullptr;
  int64_t* d_off  = nullptr;
  int      status;
  if ((status = alloc(a, &new_xu, (size_t)total * NC)) != PICNIX_OK)
    return status;
  if ((status = alloc(a, &new_gi, (size_t)total)) != PICNIX_OK)
    return status;
  if ((status = alloc(a, &d_off, (size_t)a->nseg)) != PICNIX_OK)
    return status;
  PICNIX_CUDA(a, cudaMemcpyAsync(d_off, off.data(), a->nseg * sizeof(int64_t), cudaMemcpyHostToDevice, a->stream));
  int maxcap = 0;
  for (int s = 0; s < a->nseg; s++)
    maxcap = std::max(maxcap, a->seg_cap[s]);
  const int bps = std::max(1, std::min(64, (maxcap + GROW_THREADS - 1) / GROW_THREADS));
  copy_segments_kernel<<<bps * a->nseg, GROW_THREADS, 0, a->stream>>>(a->d, d_off, new_xu, new_gi, total, bps);
  a->kernel_launches++;
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));

  cudaFree(a->d.xu);
  cudaFree(a->d.gindex);
  a->d.xu     = new_xu;
  a->d.gindex = new_gi;
  a->d.pcap   = total;
  if ((status = alloc(a, &a->d.xv, (size_t)total * NC)) != PICNIX_OK)
    return status;
  if ((status = alloc(a, &a->d.perm, (size_t)total)) != PICNIX_OK)
    return status;
  PICNIX_CUDA(a, cudaMemsetAsync(a->d.xv, 0, (size_t)total * NC * sizeof(double), a->stream));
  const int leave_cap = (int)std::min<int64_t>(std::max<int64_t>(4096, total / 4), 1 << 30);
  if (leave_cap > a->d.leave_cap) {
    cudaFree(a->d.leave_idx);
    a->d.leave_cap = leave_cap;
    if ((status = alloc(a, &a->d.leave_idx, (size_t)leave_cap)) != PICNIX_OK)
      return status;
  }
  a->leave_list_valid = false;
  a->seg_off          = off;
  a->seg_cap.assign(cap.begin(), cap.end());
  PICNIX_CUDA(a, cudaMemcpyAsync(a->d.seg_off, d_off, a->nseg * sizeof(int64_t), cudaMemcpyDeviceToDevice, a->stream));
  PICNIX_CUDA(a, cudaMemcpyAsync(a->d.seg_cap, a->seg_cap.data(), a->nseg * sizeof(int32_t),
                                 cudaMemcpyHostToDevice, a->stream));
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  cudaFree(d_off);
  a->segment_regrows++;
  return PICNIX_OK;
}

} // namespace

// Between the unpack of the particle exchange and the sort (launch_halo_end).
int resolve_growth(picnix_arena* a)
{
  if (!a->particles_allocated)
    return PICNIX_OK;
  // statistics of the previous step
  if (a->stat_pending) {
    PICNIX_CUDA(a, cudaEventSynchronize(a->stat_event)); // recorded a step ago: no stall in steady state
    a->stat_minfree = a->h_stat[0];
    a->stat_maxtail = a->h_stat[1];
    a->stat_spilled = a->h_stat[2];
    a->stat_pending = false;
    a->stat_known   = true;
    if (a->h_stat[3] != 0) // the previous step raised a fatal flag: report it now (same texts as synchronize)
      return picnix_cuda_synchronize(a);
  }
  const bool look = a->check_growth_always || !a->stat_known || a->stat_spilled > 0 ||
                    a->stat_minfree < 2 * a->stat_maxtail + 16;
  if (!look)
    return PICNIX_OK;

  const Geom&      g = a->g;
  std::vector<int> np(a->nseg), ntail(a->nseg);
  int              nspill = 0;
  PICNIX_CUDA(a, cudaMemcpyAsync(np.data(), a->d.np, a->nseg * sizeof(int), cudaMemcpyDeviceToHost, a->stream));
  PICNIX_CUDA(a, cudaMemcpyAsync(ntail.data(), a->d.ntail, a->nseg * sizeof(int), cudaMemcpyDeviceToHost, a->stream));
  PICNIX_CUDA(a, cudaMemcpyAsync(&nspill, a->d.spill_count, sizeof(int), cudaMemcpyDeviceToHost, a->stream));
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  if (nspill > a->d.spill_cap)
    return fail(a, PICNIX_ERR_OVERFLOW,
                "particle spill list overflow: more migrants than segments and spill list can hold in one step");

  // spilled migrants per destination segment
  std::vector<int> extra(a->nseg, 0);
  int              nremote = 0;
  if (nspill > 0) {
    std::vector<double> tags(nspill);
    PICNIX_CUDA(a, cudaMemcpy2D(tags.data(), sizeof(double), a->d.spill_rec + 7, 8 * sizeof(double), sizeof(double),
                                nspill, cudaMemcpyDeviceToHost));
    for (int r = 0; r < nspill; r++) {
      int2 tag;
      std::memcpy(&tag, &tags[r], sizeof(tag));
      if (tag.x >= 0 && tag.x < a->nseg)
        extra[tag.x]++;
      else
        nremote++;
    }
    if (a->stat_known && a->stat_spilled > 0)
      a->late_particles += std::min(nspill, a->stat_spilled); // they sat out the step in between
  }

  // who has to grow?  must: does not hold what is waiting; may: nearly full (avoids a look every step)
  std::vector<int32_t> cap(a->seg_cap.begin(), a->seg_cap.end());
  bool                 grow = false;
  for (int s = 0; s < a->nseg; s++) {
    const int64_t need = (int64_t)np[s] + ntail[s] + extra[s];
    if (need > a->seg_cap[s] || (extra[s] == 0 && need + 2 * (int64_t)ntail[s] + 16 > a->seg_cap[s])) {
      const int64_t want = (int64_t)((double)need * (1.0 + a->cfg.buffer_ratio)) + 2 * (int64_t)ntail[s];
      cap[s]             = (int32_t)(((want + ALLOC_UNIT) / ALLOC_UNIT) * ALLOC_UNIT);
      grow               = grow || cap[s] > a->seg_cap[s];
    }
  }
  if (grow) {
    int status = relayout(a, cap);
    if (status != PICNIX_OK)
      return status;
  }
  if (nspill > 0) {
    double* keep  = nullptr;
    int*    nkeep = nullptr;
    int     status;
    if ((status = alloc(a, &keep, (size_t)std::max(nremote, 1) * 8)) != PICNIX_OK)
      return status;
    if ((status = alloc(a, &nkeep, 1)) != PICNIX_OK)
      return status;
    PICNIX_CUDA(a, cudaMemsetAsync(nkeep, 0, sizeof(int), a->stream));
    drain_spill_kernel<<<std::min(148 * 4, (nspill + GROW_THREADS - 1) / GROW_THREADS), GROW_THREADS, 0, a->stream>>>(
        g, a->d, nspill, keep, nkeep);
    a->kernel_launches++;
    // what is left waits for the next exchange (records for remote neighbours)
    PICNIX_CUDA(a, cudaMemcpyAsync(a->d.spill_rec, keep, (size_t)nremote * 8 * sizeof(double),
                                   cudaMemcpyDeviceToDevice, a->stream));
    PICNIX_CUDA(a, cudaMemcpyAsync(a->d.spill_count, nkeep, sizeof(int), cudaMemcpyDeviceToDevice, a->stream));
    PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
    cudaFree(keep);
    cudaFree(nkeep);
  }
  return check_cuda(a, cudaGetLastError(), "resolve_growth");
}

// After the sort: statistics for the next step's decision, then the arrival counters are cleared.
int record_segment_stats(picnix_arena* a)
{
  const int init[4] = {0x7fffffff, 0, 0, 0};
  PICNIX_CUDA(a, cudaMemcpyAsync(a->d.seg_stat, init, sizeof(init), cudaMemcpyHostToDevice, a->stream));
  segment_stat_kernel<<<std::min(148, (a->nseg + 255) / 256), 256, 0, a->stream>>>(a->d, a->nseg);
  reset_tail_kernel<<<(a->nseg + 255) / 256, 256, 0, a->stream>>>(a->d, a->nseg);
  a->kernel_launches += 2;
  PICNIX_CUDA(a, cudaMemcpyAsync(a->h_stat, a->d.seg_stat, 4 * sizeof(int), cudaMemcpyDeviceToHost, a->stream));
  PICNIX_CUDA(a, cudaEventRecord(a->stat_event, a->stream));
  a->stat_pending = true;
  return check_cuda(a, cudaGetLastError(), "record_segment_stats");
}

} // namespace picnix

extern "C" int picnix_cuda_get_growth_stats(const picnix_arena_t* a, int64_t* segment_regrows, int64_t* late_particles)
{
  if (a == nullptr)
    return PICNIX_ERR_INVALID;
  if (segment_regrows)
    *segment_regrows = a->segment_regrows;
  if (late_particles)
    *late_particles = a->late_particles;
  return PICNIX_OK;
}
