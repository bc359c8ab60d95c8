// -*- C++ -*-
// Thread-per-particle kernels behind the three separate PicChunk entry points
//   push_velocity   (pic/pic_chunk.cpp:501-512 -> pic/engine/velocity.hpp)
//   push_position   (pic/pic_chunk.cpp:491-499 -> pic/engine/position.hpp + XtensorParticle::count)
//   deposit_current (pic/pic_chunk.cpp:514-523 -> pic/engine/current.hpp)
// for every (dimension, order, pusher, interpolation) combination the reference's dispatch tables
// hold (pic/pic_engine.hpp:126-195, 327-414).  They keep the reference's pass structure
// (xu/xv double buffer, separate passes); the single-pass fused kernel lives in fused.cu.
//
// Layout: particles are structure-of-arrays, so a warp's 32 loads of one component are one
// contiguous 256-B run.  Fields are read through the read-only path; a chunk's uf (<= 384 KB)
// stays L2-resident while its particles stream through.
#include "particle_kernels.cuh"

namespace picnix
{

namespace
{

constexpr int PTHREADS = 128;

struct GlobalField {
  const double* __restrict__ uf; // chunk base
  int My, Mx;
  __device__ __forceinline__ double operator()(int iz, int iy, int ix, int k) const
  {
    return __ldg(uf + ((int64_t)(iz * My + iy) * Mx + ix) * 6 + k);
  }
};

// block -> (segment, first particle); segments are (chunk, species) pairs
struct ParticleSlot {
  int     seg, chunk, is, ip;
  int64_t off;
  bool    valid;
};

__device__ __forceinline__ ParticleSlot locate(const Geom& g, const DevPtrs& d, int c0,
                                               int blocks_per_seg)
{
  ParticleSlot s;
  int          lseg = blockIdx.x / blocks_per_seg;
  int          b    = blockIdx.x - lseg * blocks_per_seg;
  s.seg   = c0 * g.Ns + lseg;
  s.chunk = s.seg / g.Ns;
  s.is    = s.seg - s.chunk * g.Ns;
  s.ip    = b * blockDim.x + threadIdx.x;
  s.off   = d.seg_off[s.seg];
  s.valid = s.ip < d.np[s.seg];
  return s;
}

template <int Dim, int Order, int Pusher, int Interp>
__global__ void __launch_bounds__(PTHREADS)
velocity_kernel(Geom g, DevPtrs d, int c0, int blocks_per_seg, double delt)
{
  ParticleSlot s = locate(g, d, c0, blocks_per_seg);
  if (!s.valid)
    return;

  const int64_t i    = s.off + s.ip;
  const double  qmdt = 0.5 * d.qm[2 * s.is] / d.qm[2 * s.is + 1] * delt;
  const double* lim  = d.clim + s.chunk * 6;
  GlobalField   F{d.uf + (int64_t)s.chunk * g.Ng * 6, g.M[1], g.M[2]};

  double x  = d.xu[0 * d.pcap + i];
  double y  = d.xu[1 * d.pcap + i];
  double z  = d.xu[2 * d.pcap + i];
  double ux = d.xu[3 * d.pcap + i];
  double uy = d.xu[4 * d.pcap + i];
  double uz = d.xu[5 * d.pcap + i];

  velocity_update<Dim, Order, Pusher, Interp>(g, lim, F, delt, qmdt, x, y, z, ux, uy, uz);

  d.xu[3 * d.pcap + i] = ux;
  d.xu[4 * d.pcap + i] = uy;
  d.xu[5 * d.pcap + i] = uz;
}

__global__ void __launch_bounds__(PTHREADS)
position_kernel(Geom g, DevPtrs d, int c0, int blocks_per_seg, double delt)
{
  ParticleSlot s = locate(g, d, c0, blocks_per_seg);
  if (!s.valid)
    return;

  const int64_t i   = s.off + s.ip;
  const double* lim = d.clim + s.chunk * 6;

  double p[NC];
#pragma unroll
  for (int k = 0; k < NC; k++) {
    p[k]                 = d.xu[k * d.pcap + i];
    d.xv[k * d.pcap + i] = p[k]; // xv <- xu, all seven components (position.hpp:120-123)
  }

  push_position(p[0], p[1], p[2], p[3], p[4], p[5], 1 / g.cc, delt);
  if (g.any_particle_bc) {
    apply_particle_bc(g, p[0], p[1], p[2], p[3], p[4], p[5]);
    d.xu[3 * d.pcap + i] = p[3];
    d.xu[4 * d.pcap + i] = p[4];
    d.xu[5 * d.pcap + i] = p[5];
  }

  d.xu[0 * d.pcap + i] = p[0];
  d.xu[1 * d.pcap + i] = p[1];
  d.xu[2 * d.pcap + i] = p[2];

  // XtensorParticle::count(0, Np-1, reset=true): the histogram was cleared by the launcher
  const int key = cell_key(g, lim, p[0], p[1], p[2]);
  d.gindex[i]   = key;
  atomicAdd(d.pcount + (int64_t)s.seg * (g.Ng + 1) + key, 1);
}

__global__ void __launch_bounds__(PTHREADS)
count_kernel(Geom g, DevPtrs d, int c0, int blocks_per_seg)
{
  ParticleSlot s = locate(g, d, c0, blocks_per_seg);
  if (!s.valid)
    return;
  const int64_t i   = s.off + s.ip;
  const double* lim = d.clim + s.chunk * 6;
  const int     key = cell_key(g, lim, d.xu[0 * d.pcap + i], d.xu[1 * d.pcap + i],
                               d.xu[2 * d.pcap + i]);
  d.gindex[i]       = key;
  atomicAdd(d.pcount + (int64_t)s.seg * (g.Ng + 1) + key, 1);
}

template <int Dim, int Order>
__global__ void __launch_bounds__(PTHREADS)
deposit_kernel(Geom g, DevPtrs d, int c0, int blocks_per_seg, double delt)
{
  ParticleSlot s = locate(g, d, c0, blocks_per_seg);
  if (!s.valid)
    return;

  const int64_t i   = s.off + s.ip;
  const double* lim = d.clim + s.chunk * 6;
  const double  q   = d.qm[2 * s.is];
  double*       uj  = d.uj + (int64_t)s.chunk * g.Ng * 4;

  // before: xv, after: xu
  const double x0 = d.xv[0 * d.pcap + i], y0 = d.xv[1 * d.pcap + i], z0 = d.xv[2 * d.pcap + i];
  const double x1 = d.xu[0 * d.pcap + i], y1 = d.xu[1 * d.pcap + i], z1 = d.xu[2 * d.pcap + i];

  // stencil base; assigned by esirkepov_deposit before its first call to add()
  int bz = 0, by = 0, bx = 0;
  const int My = g.M[1], Mx = g.M[2];
  auto add = [&](int jz, int jy, int jx, int k, double v) {
    if (v != 0.0) {
      atomicAdd(uj + ((int64_t)((bz + jz) * My + (by + jy)) * Mx + (bx + jx)) * 4 + k, v);
    }
  };
  esirkepov_deposit<Dim, Order>(g, lim, q, delt, x0, y0, z0, x1, y1, z1, bz, by, bx, add);
}

int blocks_per_segment(const picnix_arena* a, int c0, int cn)
{
  int maxcap = 0;
  for (int s = c0 * a->g.Ns; s < (c0 + cn) * a->g.Ns; s++)
    maxcap = std::max(maxcap, a->seg_cap[s]);
  return (maxcap + PTHREADS - 1) / PTHREADS;
}

template <int Dim, int Order, int Pusher>
void launch_velocity_interp(picnix_arena* a, int c0, int cn, int bps, double delt)
{
  int blocks = bps * cn * a->g.Ns;
  if (a->cfg.interp == PICNIX_INTERP_MC) {
    velocity_kernel<Dim, Order, Pusher, PICNIX_INTERP_MC>
        <<<blocks, PTHREADS, 0, a->stream>>>(a->g, a->d, c0, bps, delt);
  } else {
    velocity_kernel<Dim, Order, Pusher, PICNIX_INTERP_WT>
        <<<blocks, PTHREADS, 0, a->stream>>>(a->g, a->d, c0, bps, delt);
  }
}

template <int Dim, int Order>
void launch_velocity_pusher(picnix_arena* a, int c0, int cn, int bps, double delt)
{
  switch (a->cfg.pusher) {
  case PICNIX_PUSHER_BORIS:
    launch_velocity_interp<Dim, Order, PICNIX_PUSHER_BORIS>(a, c0, cn, bps, delt);
    break;
  case PICNIX_PUSHER_VAY:
    launch_velocity_interp<Dim, Order, PICNIX_PUSHER_VAY>(a, c0, cn, bps, delt);
    break;
  default:
    launch_velocity_interp<Dim, Order, PICNIX_PUSHER_HIGUERA_CARY>(a, c0, cn, bps, delt);
    break;
  }
}

template <int Dim>
void launch_velocity_order(picnix_arena* a, int c0, int cn, int bps, double delt)
{
  switch (a->g.order) {
  case 1:
    launch_velocity_pusher<Dim, 1>(a, c0, cn, bps, delt);
    break;
  case 2:
    launch_velocity_pusher<Dim, 2>(a, c0, cn, bps, delt);
    break;
  case 3:
    launch_velocity_pusher<Dim, 3>(a, c0, cn, bps, delt);
    break;
  default:
    launch_velocity_pusher<Dim, 4>(a, c0, cn, bps, delt);
    break;
  }
}

template <int Dim>
void launch_deposit_order(picnix_arena* a, int c0, int cn, int bps, double delt)
{
  int blocks = bps * cn * a->g.Ns;
  switch (a->g.order) {
  case 1:
    deposit_kernel<Dim, 1><<<blocks, PTHREADS, 0, a->stream>>>(a->g, a->d, c0, bps, delt);
    break;
  case 2:
    deposit_kernel<Dim, 2><<<blocks, PTHREADS, 0, a->stream>>>(a->g, a->d, c0, bps, delt);
    break;
  case 3:
    deposit_kernel<Dim, 3><<<blocks, PTHREADS, 0, a->stream>>>(a->g, a->d, c0, bps, delt);
    break;
  default:
    deposit_kernel<Dim, 4><<<blocks, PTHREADS, 0, a->stream>>>(a->g, a->d, c0, bps, delt);
    break;
  }
}

} // namespace

int launch_push_velocity(picnix_arena* a, int c0, int cn, double delt)
{
  resolve_range(a, c0, cn);
  {
    int mstatus = materialize_sort(a); // a pending index-only sort must be made physical first
    if (mstatus != PICNIX_OK)
      return mstatus;
  }
  if (!a->particles_allocated)
    return fail(a, PICNIX_ERR_INVALID, "no particles allocated");
  int bps = blocks_per_segment(a, c0, cn);
  if (bps == 0 || cn == 0)
    return PICNIX_OK;
  switch (a->g.dimension) {
  case 1:
    launch_velocity_order<1>(a, c0, cn, bps, delt);
    break;
  case 2:
    launch_velocity_order<2>(a, c0, cn, bps, delt);
    break;
  default:
    launch_velocity_order<3>(a, c0, cn, bps, delt);
    break;
  }
  a->kernel_launches++;
  return check_cuda(a, cudaGetLastError(), "push_velocity");
}

int launch_push_position(picnix_arena* a, int c0, int cn, double delt)
{
  resolve_range(a, c0, cn);
  {
    int mstatus = materialize_sort(a); // a pending index-only sort must be made physical first
    if (mstatus != PICNIX_OK)
      return mstatus;
  }
  if (!a->particles_allocated)
    return fail(a, PICNIX_ERR_INVALID, "no particles allocated");
  int bps = blocks_per_segment(a, c0, cn);
  if (bps == 0 || cn == 0)
    return PICNIX_OK;
  const int64_t nbin = a->g.Ng + 1;
  PICNIX_CUDA(a, cudaMemsetAsync(a->d.pcount + (int64_t)c0 * a->g.Ns * nbin, 0,
                                 (size_t)cn * a->g.Ns * nbin * sizeof(int), a->stream));
  a->leave_list_valid = false;
  position_kernel<<<bps * cn * a->g.Ns, PTHREADS, 0, a->stream>>>(a->g, a->d, c0, bps, delt);
  a->kernel_launches++;
  return check_cuda(a, cudaGetLastError(), "push_position");
}

int launch_count(picnix_arena* a, int c0, int cn)
{
  resolve_range(a, c0, cn);
  {
    int mstatus = materialize_sort(a); // a pending index-only sort must be made physical first
    if (mstatus != PICNIX_OK)
      return mstatus;
  }
  if (!a->particles_allocated)
    return fail(a, PICNIX_ERR_INVALID, "no particles allocated");
  int bps = blocks_per_segment(a, c0, cn);
  if (bps == 0 || cn == 0)
    return PICNIX_OK;
  const int64_t nbin = a->g.Ng + 1;
  PICNIX_CUDA(a, cudaMemsetAsync(a->d.pcount + (int64_t)c0 * a->g.Ns * nbin, 0,
                                 (size_t)cn * a->g.Ns * nbin * sizeof(int), a->stream));
  a->leave_list_valid = false;
  count_kernel<<<bps * cn * a->g.Ns, PTHREADS, 0, a->stream>>>(a->g, a->d, c0, bps);
  a->kernel_launches++;
  return check_cuda(a, cudaGetLastError(), "count");
}

int launch_deposit_current(picnix_arena* a, int c0, int cn, double delt)
{
  resolve_range(a, c0, cn);
  {
    int mstatus = materialize_sort(a); // a pending index-only sort must be made physical first
    if (mstatus != PICNIX_OK)
      return mstatus;
  }
  if (!a->particles_allocated)
    return fail(a, PICNIX_ERR_INVALID, "no particles allocated");
  if (cn == 0)
    return PICNIX_OK;
  // fill_all(uj, 0), pic/engine/current.hpp:91,152
  PICNIX_CUDA(a, cudaMemsetAsync(a->d.uj + (int64_t)c0 * a->g.Ng * 4, 0,
                                 (size_t)cn * a->g.Ng * 4 * sizeof(double), a->stream));
  if (row_kernel_applies(a))
    return launch_deposit_rows(a, c0, cn, delt);
  int bps = blocks_per_segment(a, c0, cn);
  if (bps == 0)
    return PICNIX_OK;
  switch (a->g.dimension) {
  case 1:
    launch_deposit_order<1>(a, c0, cn, bps, delt);
    break;
  case 2:
    launch_deposit_order<2>(a, c0, cn, bps, delt);
    break;
  default:
    launch_deposit_order<3>(a, c0, cn, bps, delt);
    break;
  }
  a->kernel_launches++;
  return check_cuda(a, cudaGetLastError(), "deposit_current");
}

} // namespace picnix
