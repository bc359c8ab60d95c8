// -*- C++ -*-
// Fused K1+K2: field interpolation + momentum push + position push + cell key + Esirkepov deposit
// in ONE pass over the particles.
//
// The reference makes three passes over the particle arrays per step (push_velocity,
// push_position(+count), deposit_current; pic/pic_application.cpp:236-240) and materialises the old
// state in xv (pic/engine/position.hpp:120-123).  Here the old position stays in registers, the new
// state overwrites xu in place and the cell key + histogram (XtensorParticle::count) are produced
// on the way, so the pass reads 56 B and writes 56 B + 4 B per particle.
//
// Generic variant (any dimension / order / pusher / interpolation): one thread per particle,
// fields through the read-only path, currents through fp64 global reductions (RED.ADD.F64).
// The tiled shared-memory variant for 3-D lives in rowfused.cu and is selected when it applies.
#include "particle_kernels.cuh"

namespace picnix
{

namespace
{

constexpr int FTHREADS = 128;

struct GlobalFieldF {
  const double* __restrict__ uf;
  int My, Mx;
  __device__ __forceinline__ double operator()(int iz, int iy, int ix, int k) const
  {
    return __ldg(uf + ((int64_t)(iz * My + iy) * Mx + ix) * 6 + k);
  }
};

template <int Dim, int Order, int Pusher, int Interp>
__global__ void __launch_bounds__(FTHREADS)
fused_generic_kernel(Geom g, DevPtrs d, int c0, int blocks_per_seg, double delt)
{
  const int lseg  = blockIdx.x / blocks_per_seg;
  const int b     = blockIdx.x - lseg * blocks_per_seg;
  const int seg   = c0 * g.Ns + lseg;
  const int chunk = seg / g.Ns;
  const int is    = seg - chunk * g.Ns;
  const int ip    = b * blockDim.x + threadIdx.x;
  if (ip >= d.np[seg])
    return;

  const int64_t i    = d.seg_off[seg] + ip;
  const double  q    = d.qm[2 * is];
  const double  qmdt = 0.5 * q / d.qm[2 * is + 1] * delt;
  const double* lim  = d.clim + chunk * 6;
  GlobalFieldF  F{d.uf + (int64_t)chunk * g.Ng * 6, g.M[1], g.M[2]};

  const double x0 = d.xu[0 * d.pcap + i];
  const double y0 = d.xu[1 * d.pcap + i];
  const double z0 = d.xu[2 * d.pcap + i];
  double       ux = d.xu[3 * d.pcap + i];
  double       uy = d.xu[4 * d.pcap + i];
  double       uz = d.xu[5 * d.pcap + i];

  velocity_update<Dim, Order, Pusher, Interp>(g, lim, F, delt, qmdt, x0, y0, z0, ux, uy, uz);

  double x1 = x0, y1 = y0, z1 = z0;
  push_position(x1, y1, z1, ux, uy, uz, 1 / g.cc, delt);
  apply_particle_bc(g, x1, y1, z1, ux, uy, uz);

  d.xu[0 * d.pcap + i] = x1;
  d.xu[1 * d.pcap + i] = y1;
  d.xu[2 * d.pcap + i] = z1;
  d.xu[3 * d.pcap + i] = ux;
  d.xu[4 * d.pcap + i] = uy;
  d.xu[5 * d.pcap + i] = uz;

  const int key = cell_key(g, lim, x1, y1, z1);
  d.gindex[i]   = key;
  atomicAdd(d.pcount + (int64_t)seg * (g.Ng + 1) + key, 1);
  if (key == g.Ng)
    note_leaver(d, seg, i);

  double*   uj = d.uj + (int64_t)chunk * g.Ng * 4;
  int       bz = 0, by = 0, bx = 0;
  const int My = g.M[1], Mx = g.M[2];
  auto      add = [&](int jz, int jy, int jx, int k, double v) {
    if (v != 0.0) {
      atomicAdd(uj + ((int64_t)((bz + jz) * My + (by + jy)) * Mx + (bx + jx)) * 4 + k, v);
    }
  };
  esirkepov_deposit<Dim, Order>(g, lim, q, delt, x0, y0, z0, x1, y1, z1, bz, by, bx, add);
}


template <int Dim, int Order, int Pusher>
void launch_generic_interp(picnix_arena* a, int c0, int blocks, int bps, double delt)
{
  if (a->cfg.interp == PICNIX_INTERP_MC) {
    fused_generic_kernel<Dim, Order, Pusher, PICNIX_INTERP_MC>
        <<<blocks, FTHREADS, 0, a->stream>>>(a->g, a->d, c0, bps, delt);
  } else {
    fused_generic_kernel<Dim, Order, Pusher, PICNIX_INTERP_WT>
        <<<blocks, FTHREADS, 0, a->stream>>>(a->g, a->d, c0, bps, delt);
  }
}

template <int Dim, int Order>
void launch_generic_pusher(picnix_arena* a, int c0, int blocks, int bps, double delt)
{
  switch (a->cfg.pusher) {
  case PICNIX_PUSHER_BORIS:
    launch_generic_interp<Dim, Order, PICNIX_PUSHER_BORIS>(a, c0, blocks, bps, delt);
    break;
  case PICNIX_PUSHER_VAY:
    launch_generic_interp<Dim, Order, PICNIX_PUSHER_VAY>(a, c0, blocks, bps, delt);
    break;
  default:
    launch_generic_interp<Dim, Order, PICNIX_PUSHER_HIGUERA_CARY>(a, c0, blocks, bps, delt);
    break;
  }
}

template <int Dim>
void launch_generic_order(picnix_arena* a, int c0, int blocks, int bps, double delt)
{
  switch (a->g.order) {
  case 1:
    launch_generic_pusher<Dim, 1>(a, c0, blocks, bps, delt);
    break;
  case 2:
    launch_generic_pusher<Dim, 2>(a, c0, blocks, bps, delt);
    break;
  case 3:
    launch_generic_pusher<Dim, 3>(a, c0, blocks, bps, delt);
    break;
  default:
    launch_generic_pusher<Dim, 4>(a, c0, blocks, bps, delt);
    break;
  }
}

} // namespace

int launch_row_fused(picnix_arena* a, int c0, int cn, double delt); // rowfused.cu

int launch_push_deposit_fused(picnix_arena* a, int c0, int cn, double delt)
{
  resolve_range(a, c0, cn);
  if (!a->particles_allocated)
    return fail(a, PICNIX_ERR_INVALID, "no particles allocated");
  if (cn == 0)
    return PICNIX_OK;

  const Geom&   g    = a->g;
  const int64_t nbin = g.Ng + 1;
  // fill_all(uj, 0) and XtensorParticle::reset_count
  PICNIX_CUDA(a, cudaMemsetAsync(a->d.uj + (int64_t)c0 * g.Ng * 4, 0,
                                 (size_t)cn * g.Ng * 4 * sizeof(double), a->stream));
  PICNIX_CUDA(a, cudaMemsetAsync(a->d.pcount + (int64_t)c0 * g.Ns * nbin, 0,
                                 (size_t)cn * g.Ns * nbin * sizeof(int), a->stream));

  // the kernels below list the particles that leave their chunk for the migration
  PICNIX_CUDA(a, cudaMemsetAsync(a->d.leave_count, 0, sizeof(int), a->stream));
  a->leave_list_valid = (c0 == 0 && cn == g.nchunk);

  if (row_kernel_applies(a))
    return launch_row_fused(a, c0, cn, delt); // consumes a pending index-only sort itself

  {
    int mstatus = materialize_sort(a); // a pending index-only sort must be made physical first
    if (mstatus != PICNIX_OK)
      return mstatus;
  }

  int maxcap = 0;
  for (int s = c0 * g.Ns; s < (c0 + cn) * g.Ns; s++)
    maxcap = std::max(maxcap, a->seg_cap[s]);
  int bps = (maxcap + FTHREADS - 1) / FTHREADS;
  if (bps == 0)
    return PICNIX_OK;
  int blocks = bps * cn * g.Ns;

  switch (g.dimension) {
  case 1:
    launch_generic_order<1>(a, c0, blocks, bps, delt);
    break;
  case 2:
    launch_generic_order<2>(a, c0, blocks, bps, delt);
    break;
  default:
    launch_generic_order<3>(a, c0, blocks, bps, delt);
    break;
  }
  a->kernel_launches++;
  return check_cuda(a, cudaGetLastError(), "push_deposit_fused");
}

} // namespace picnix
