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
// The tiled shared-memory variant for 3-D lives below it and is selected when it applies.
#include "particle_kernels.cuh"
#include "rowdeposit.cuh"

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

  d.xu[0 * d.pcap + i] = x1;
  d.xu[1 * d.pcap + i] = y1;
  d.xu[2 * d.pcap + i] = z1;
  d.xu[3 * d.pcap + i] = ux;
  d.xu[4 * d.pcap + i] = uy;
  d.xu[5 * d.pcap + i] = uz;

  const int key = cell_key(g, lim, x1, y1, z1);
  d.gindex[i]   = key;
  atomicAdd(d.pcount + (int64_t)seg * (g.Ng + 1) + key, 1);

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


// ---------------------------------------------------------------------------------------------
// Row-owner kernel (3-D, order 2): see rowdeposit.cuh for the scheme.
//   FUSED = true : phase 1 also interpolates the fields, pushes momentum and position, writes the
//                  new state in place and produces the cell key + histogram (K1 + K2 in one pass)
//   FUSED = false: deposit only, old position from xv, new position from xu (PicChunk::deposit_current)
// One warp per (chunk, z, y, x-segment of RX cells); the particles of the segment are contiguous in
// the cell-sorted arrays: [pindex[key0], pindex[key0 + RX]).
// ---------------------------------------------------------------------------------------------
template <bool FUSED, int Pusher, int Interp>
__global__ void __launch_bounds__(rowdep::THREADS, 2)
row_kernel(Geom g, DevPtrs d, int c0, int cn, double delt)
{
  using namespace rowdep;
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const int      lane = threadIdx.x & 31;
  const int      warp = threadIdx.x >> 5;
  const int      half = lane >> 4;
  const int      a    = (lane >> 2) & 3;
  const int      b    = lane & 3;
  const unsigned FULL = 0xffffffffu;
  WarpSmem*      ws   = reinterpret_cast<WarpSmem*>(smem_raw) + warp;

  const int nsegx = g.dims[2] / RX;
  const int rows  = g.dims[0] * g.dims[1] * nsegx;
  const int gw    = blockIdx.x * WARPS + warp;
  const int lc    = gw / rows;
  if (lc >= cn)
    return;
  int       r     = gw - lc * rows;
  const int jz    = r / (g.dims[1] * nsegx);
  r -= jz * g.dims[1] * nsegx;
  const int jy    = r / nsegx;
  const int jx0   = (r - jy * nsegx) * RX;
  const int chunk = c0 + lc;

  const double* lim = d.clim + chunk * 6;
  double*       uj  = d.uj + (int64_t)chunk * g.Ng * 4;
  const int     My = g.M[1], Mx = g.M[2];

  for (int i = lane; i < TILE; i += 32)
    ws->tile[i] = 0.0;
  __syncwarp();

  const double dxdt = g.del[2] / delt, dydt = g.del[1] / delt, dzdt = g.del[0] / delt;
  const int    key0 = jz * g.fsz + jy * g.fsy + jx0;

  for (int is = 0; is < g.Ns; is++) {
    const int     seg = chunk * g.Ns + is;
    const int64_t off = d.seg_off[seg];
    const int*    pix = d.pindex + (int64_t)seg * (g.Ng + 1);
    const int     pb = pix[key0], pe = pix[key0 + RX];
    const double  q    = d.qm[2 * is];
    const double  qmdt = 0.5 * q / d.qm[2 * is + 1] * delt;

    Acc acc;
    acc.clear();
    int cur = -1;

    for (int base = pb; base < pe; base += 32) {
      const int n = min(32, pe - base);

      // ---------------- phase 1: one particle per lane ----------------
      int inf = 0;
      if (lane < n) {
        const int64_t i = off + base + lane;
        double        x0, y0, z0, x1, y1, z1;
        if (FUSED) {
          x0        = d.xu[0 * d.pcap + i];
          y0        = d.xu[1 * d.pcap + i];
          z0        = d.xu[2 * d.pcap + i];
          double ux = d.xu[3 * d.pcap + i];
          double uy = d.xu[4 * d.pcap + i];
          double uz = d.xu[5 * d.pcap + i];
          GlobalFieldF F{d.uf + (int64_t)chunk * g.Ng * 6, My, Mx};
          velocity_update<3, 2, Pusher, Interp>(g, lim, F, delt, qmdt, x0, y0, z0, ux, uy, uz);
          x1 = x0;
          y1 = y0;
          z1 = z0;
          push_position(x1, y1, z1, ux, uy, uz, 1 / g.cc, delt);
          d.xu[0 * d.pcap + i] = x1;
          d.xu[1 * d.pcap + i] = y1;
          d.xu[2 * d.pcap + i] = z1;
          d.xu[3 * d.pcap + i] = ux;
          d.xu[4 * d.pcap + i] = uy;
          d.xu[5 * d.pcap + i] = uz;
          const int key = cell_key(g, lim, x1, y1, z1);
          d.gindex[i]   = key;
          atomicAdd(d.pcount + (int64_t)seg * (g.Ng + 1) + key, 1);
        } else {
          x0 = d.xv[0 * d.pcap + i];
          y0 = d.xv[1 * d.pcap + i];
          z0 = d.xv[2 * d.pcap + i];
          x1 = d.xu[0 * d.pcap + i];
          y1 = d.xu[1 * d.pcap + i];
          z1 = d.xu[2 * d.pcap + i];
        }

        AxisFactors fx = axis_factors(x0, x1, lim[4], g.del[2]);
        AxisFactors fy = axis_factors(y0, y1, lim[2], g.del[1]);
        AxisFactors fz = axis_factors(z0, z1, lim[0], g.del[0]);
        const int   jx = fx.i0 - jx0;
        if (fx.ok && fy.ok && fz.ok && fy.i0 == jy && fz.i0 == jz && jx >= 0 && jx < RX) {
          stage_particle(ws->stg + lane * NSTG, fx, fy, fz, q, dxdt, dydt, dzdt);
          inf = make_info(jx, fx.w, fy.w, fz.w);
        } else {
          // not where the sort says it is (or moved more than a cell): generic atomic deposit
          int  bz = 0, by = 0, bx = 0;
          auto add = [&](int kz, int ky, int kx, int k, double v) {
            if (v != 0.0)
              atomicAdd(uj + ((int64_t)((bz + kz) * My + (by + ky)) * Mx + (bx + kx)) * 4 + k, v);
          };
          esirkepov_deposit<3, 2>(g, lim, q, delt, x0, y0, z0, x1, y1, z1, bz, by, bx, add);
        }
      }
      ws->info[lane] = inf;
      __syncwarp();

      // ---------------- phase 2: one staged particle per half-warp ----------------
      const int npass = (n + 1) >> 1;
      for (int k = 0; k < npass; k++) {
        const int     j     = 2 * k + half;
        const int     pinf  = j < n ? ws->info[j] : 0;
        const bool    valid = (pinf >> 11) & 1;
        const int     jx    = pinf & 0xff;
        const int     wx = (pinf >> 8) & 1, wy = (pinf >> 9) & 1, wz = (pinf >> 10) & 1;
        const bool    major = valid && (wx & wy & wz);
        const bool    minor = valid && !major;
        const double* rec   = ws->stg + j * NSTG;

        const bool newcell = major && cur >= 0 && cur != jx;
        if (__any_sync(FULL, newcell)) {
#pragma unroll
          for (int hh = 0; hh < 2; hh++) {
            if (half == hh && newcell)
              flush(ws->tile, acc, a, b, cur, 1, 1, 1);
            __syncwarp();
          }
          if (newcell)
            acc.clear();
        }
        if (major) {
          cur = jx;
          accumulate(acc, rec, a, b);
        }
        if (__any_sync(FULL, minor)) {
          Acc tmp;
          tmp.clear();
          if (minor)
            accumulate(tmp, rec, a, b);
#pragma unroll
          for (int hh = 0; hh < 2; hh++) {
            if (half == hh && minor)
              flush(ws->tile, tmp, a, b, jx, wx, wy, wz);
            __syncwarp();
          }
        }
      }
      __syncwarp();
    }

    // end of this species' particles in the segment
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      if (half == hh && cur >= 0)
        flush(ws->tile, acc, a, b, cur, 1, 1, 1);
      __syncwarp();
    }
  }

  // warp tile -> global current: one fp64 reduction per non-zero tile value
  const int gz0 = jz + g.Lb[0] - 2, gy0 = jy + g.Lb[1] - 2, gx0 = jx0 + g.Lb[2] - 2;
  for (int idx = lane; idx < 25 * XS * 4; idx += 32) {
    const int    tz = idx / (5 * XS * 4);
    const int    r2 = idx - tz * (5 * XS * 4);
    const int    ty = r2 / (XS * 4);
    const int    e  = r2 - ty * (XS * 4);
    const double v  = ws->tile[tz * SZ + ty * SY + e];
    if (v != 0.0) {
      atomicAdd(uj + ((int64_t)((gz0 + tz) * My + (gy0 + ty)) * Mx + gx0) * 4 + e, v);
    }
  }
}

template <bool FUSED>
int launch_row_kernel(picnix_arena* a, int c0, int cn, double delt)
{
  using namespace rowdep;
  const Geom& g      = a->g;
  const int   rows   = g.dims[0] * g.dims[1] * (g.dims[2] / RX);
  const int   blocks = (rows * cn + WARPS - 1) / WARPS;
  const int   key    = FUSED ? a->cfg.pusher * 2 + a->cfg.interp : 0;

#define PICNIX_ROW_LAUNCH(P, I)                                                                    \
  {                                                                                                \
    auto kern = row_kernel<FUSED, P, I>;                                                           \
    PICNIX_CUDA(a, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                        (int)SMEM_BYTES));                                         \
    kern<<<blocks, THREADS, SMEM_BYTES, a->stream>>>(g, a->d, c0, cn, delt);                       \
  }
  switch (key) {
  case 0:
    PICNIX_ROW_LAUNCH(PICNIX_PUSHER_BORIS, PICNIX_INTERP_MC);
    break;
  case 1:
    PICNIX_ROW_LAUNCH(PICNIX_PUSHER_BORIS, PICNIX_INTERP_WT);
    break;
  case 2:
    PICNIX_ROW_LAUNCH(PICNIX_PUSHER_VAY, PICNIX_INTERP_MC);
    break;
  case 3:
    PICNIX_ROW_LAUNCH(PICNIX_PUSHER_VAY, PICNIX_INTERP_WT);
    break;
  case 4:
    PICNIX_ROW_LAUNCH(PICNIX_PUSHER_HIGUERA_CARY, PICNIX_INTERP_MC);
    break;
  default:
    PICNIX_ROW_LAUNCH(PICNIX_PUSHER_HIGUERA_CARY, PICNIX_INTERP_WT);
    break;
  }
#undef PICNIX_ROW_LAUNCH
  a->kernel_launches++;
  return check_cuda(a, cudaGetLastError(), "row_kernel");
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

// The row-owner kernel needs 3-D, 2nd-order shapes, rows that split into RX-cell segments and a
// pindex that describes the current particle order (set by the sort, cleared by uploads).
bool row_kernel_applies(const picnix_arena* a)
{
  const Geom& g = a->g;
  return g.dimension == 3 && g.order == 2 && (g.dims[2] % rowdep::RX) == 0 && a->pindex_valid &&
         !a->force_generic;
}

int launch_deposit_rows(picnix_arena* a, int c0, int cn, double delt)
{
  return launch_row_kernel<false>(a, c0, cn, delt);
}

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

  if (row_kernel_applies(a))
    return launch_row_kernel<true>(a, c0, cn, delt);

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
