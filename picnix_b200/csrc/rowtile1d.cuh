// -*- C++ -*-
// Tiled push + Esirkepov deposit for 1-D runs (x; y and z ignorable) with 2nd-order shapes: the kernel of
// rowtile.cuh / rowpush.cu for the BASELINE configurations that live in one dimension (two-stream, shock).
// Same construction -- one merged particle stream per 8-cell segment with cells aligned to even slots,
// cell-anchored interpolation, staged records consumed by half-warps, warp-private current tile, one fp64
// reduction per non-zero tile value -- with the 1-D form of the density decomposition
// (nix/esirkepov.hpp:18-74 on the 4-slot window; pic/engine/current.hpp:220-221 for the ignorable axes):
//     rho[x]   += q S1x[x]
//     Jx [x+1] += -q dx/dt * prefix sums of DSx
//     Jy [x]   += q vy (S0x[x] + DSx[x]/2)          vy = (y_new - y_old) / dt
//     Jz [x]   += q vz (S0x[x] + DSx[x]/2)          vz = (z_new - z_old) / dt
// The sixteen values of a particle are formed by its own lane in phase 1; phase 2 is a sum over the
// particles of a cell: lane (c, i) of a half-warp adds value 4 c + i of the half-warp's record.
#ifndef PICNIX_B200_ROWTILE1D_CUH
#define PICNIX_B200_ROWTILE1D_CUH

#include "rowtile.cuh"

namespace picnix
{
namespace rowtile1d
{

using rowtile::ALIGN;
using rowtile::AxisFactors;
using rowtile::cp_async_16;
using rowtile::cp_async_commit_wait;
using rowtile::cp_async_f64;
using rowtile::cp_async_i32;
using rowtile::make_info;
using rowtile::MAXNS;
using rowtile::push_boris_fast;
using rowtile::push_position_fast;
using rowtile::RX;
using rowtile::shape2;
using rowtile::shift4;
using rowtile::store2;
using rowtile::THREADS;
using rowtile::WARPS;
using rowtile::window_factors;

// field tile of a warp: points x in [jx0-1, jx0+RX+1] of the one (z, y) line, layout [x][6]
constexpr int FX    = RX + 3;
constexpr int FTILE = FX * 6;

// current tile of a warp: [x RX+4][component 4]; element 4 * x + c
constexpr int XS   = RX + 4;
constexpr int TILE = 4 * XS;

// staged record (doubles): rho[4] | Jx[3], 0 | Jy[4] | Jz[4], padded to an odd number of 16-byte words
constexpr int REC = 18;

struct WarpSmem {
  double  stg[32 * REC];
  double  tile[TILE];
  double  ftile[FTILE];
  double  pfb[7][32];                // phase space of the next batch, filled by cp.async
  double  zero[REC];                 // the all-zero record
  double  rowc[4];                   // chunk limits and first grid point
  int64_t off[MAXNS];                // first slot of the (chunk, species) segments of this warp's chunk
  int     info[32];
  int     pbuf[32];
  int4    ent[MAXNS * RX + 1];       // merged stream, as in rowtile.cuh
};
static_assert(sizeof(WarpSmem) % 16 == 0, "the records of the next warp must stay 16-byte aligned");

struct BlockSmem {
  double q[MAXNS];                   // charge
  double qmdt[MAXNS];                // q/m dt/2
};

constexpr size_t SMEM_BYTES = sizeof(BlockSmem) + sizeof(WarpSmem) * WARPS;

template <int NX>
__device__ __forceinline__ double interp_cell(const double* __restrict__ p, const double* wx)
{
  double rx = 0;
#pragma unroll
  for (int jx = 0; jx < NX; jx++)
    rx += p[jx * 6] * wx[jx];
  return rx;
}

// phase 1: the sixteen contributions of one particle on its 4-slot window
__device__ __forceinline__ void stage_particle(double* __restrict__ rec, const AxisFactors& fx, double q,
                                               double qvy, double qvz, double dxdt)
{
  const double A  = 1.0 / 2;
  const double cx = -q * dxdt;
  double       ax[4];
#pragma unroll
  for (int k = 0; k < 4; k++)
    ax[k] = fx.S0[k] + A * fx.DS[k];
  const double px0 = fx.DS[0], px1 = px0 + fx.DS[1], px2 = px1 + fx.DS[2];
  store2(rec + 0, q * fx.S1[0], q * fx.S1[1]);
  store2(rec + 2, q * fx.S1[2], q * fx.S1[3]);
  store2(rec + 4, cx * px0, cx * px1);
  store2(rec + 6, cx * px2, 0.0);
  store2(rec + 8, qvy * ax[0], qvy * ax[1]);
  store2(rec + 10, qvy * ax[2], qvy * ax[3]);
  store2(rec + 12, qvz * ax[0], qvz * ax[1]);
  store2(rec + 14, qvz * ax[2], qvz * ax[3]);
}

// lane (c, i) = (component, window slot) inside its half-warp
struct LaneMap {
  int v;    // value of the record this lane sums
  int lin;  // lane part of the tile element index (the run adds 4 * (jx + wx))
  bool real;
};

__device__ __forceinline__ LaneMap lane_map(int lane)
{
  const int i = lane & 3;
  const int c = (lane >> 2) & 3;
  LaneMap   m;
  m.v    = 4 * c + i;
  m.lin  = 4 * (i + (c == 1 ? 1 : 0)) + c; // Jx lives one point to the right of its prefix sum
  m.real = c != 1 || i < 3;
  return m;
}

// both half-warps hold a partial sum of the same cell
__device__ __forceinline__ void flush(double* __restrict__ tile, double acc, const LaneMap& m, int run, int half)
{
  const double sum = acc + __shfl_xor_sync(0xffffffffu, acc, 16);
  if (half == 0 && m.real)
    tile[4 * run + m.lin] += sum;
}

// one staged particle straight into the tile
__device__ __forceinline__ void deposit_direct(double* __restrict__ tile, const double* __restrict__ rec,
                                               const LaneMap& m, int run, int half)
{
  if (half == 0 && m.real)
    tile[4 * run + m.lin] += rec[m.v];
}

__device__ __forceinline__ int run_index(int info)
{
  return (info & 0xff) + ((info >> 8) & 1);
}

} // namespace rowtile1d
} // namespace picnix

#endif
