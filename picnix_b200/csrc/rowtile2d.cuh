// -*- C++ -*-
// Row-owner push + Esirkepov deposit for 2-D runs (x, y; z ignorable) with 2nd-order shapes: the tiled kernel
// of rowtile.cuh / rowpush.cu for the BASELINE configurations that live in two dimensions (cherenkov, mrx).
// Same construction -- one merged particle stream per row segment with cells aligned to even slots,
// cell-anchored interpolation, staged factor records consumed by half-warps, warp-private current tile,
// one fp64 reduction per non-zero tile value -- with the 2-D form of the density decomposition
// (nix/esirkepov.hpp:76-141 on the 4 x 4 window):
//     rho[y][x]   += S1y[y] * (q S1x[x])
//     Jx [y][x+1] += (S0y[y] + DSy[y]/2) * Px[x]                 Px = -q dx/dt * prefix sums of DSx
//     Jy [y+1][x] += (S0x[x] + DSx[x]/2) * Py[y]                 Py = -q dy/dt * prefix sums of DSy
//     Jz [y][x]   += S0y[y] * (q vz AX[x]) + DSy[y] * (q vz BX[x])   AX = S0x + DSx/2, BX = S0x/2 + DSx/3,
//                                                                 vz = (z_new - z_old) / dt
// Lane (c, a) of a half-warp owns component c and ONE line of the window: row y = a for rho, Jx, Jz, column
// x = a for Jy; its four accumulators follow acc[j] += P * R1[j] + Q * R2[j] with lane-constant table
// offsets (Q is the constant 0 except for Jz).
#ifndef PICNIX_B200_ROWTILE2D_CUH
#define PICNIX_B200_ROWTILE2D_CUH

#include "rowtile.cuh"

namespace picnix
{
namespace rowtile2d
{

using rowtile::ALIGN;
using rowtile::AxisFactors;
using rowtile::BlockSmem;
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

// field tile: points x in [jx0-1, jx0+RX+1], y in [jy0-1, jy0+WARPS+1] of the one z plane, layout [y][x][6]
constexpr int FX    = RX + 3;
constexpr int FY    = WARPS + 3;
constexpr int FROW  = FX * 6;
constexpr int FTILE = FY * FROW;

// current tile of a warp: [y 5][x RX+4 (+1 pad)][component 4]; element 4 * lin + c, lin = y * SY + x.
// SY is odd so that the lanes a = 0..3 of the row-owning components fall on different banks (Jy owns
// columns: stride 1).
constexpr int XS   = RX + 4;
constexpr int SY   = XS + 1;
constexpr int TILE = 4 * 5 * SY;

// staged record (doubles): P[c][a] | Q: DSy[a], 0 | R1[c][j] | R2: q vz BX[j]
constexpr int REC  = 42;             // 21 x 16 B: odd multiple -> conflict-free 128-bit stores
constexpr int T_P  = 0;              // S1y[4] | WY[4] | AX[4] | S0y[4]
constexpr int T_Q  = 16;             // DSy[4], then the constant 0 at T_Q + 4
constexpr int T_R1 = 22;             // q S1x[4] | Px[3], 0 | Py[3], 0 | q vz AX[4]
constexpr int T_R2 = 38;             // q vz BX[4]

struct WarpSmem {
  double stg[32 * REC];
  double tile[TILE];
  double pfb[7][32];                 // phase space of the next batch, filled by cp.async
  double zero[REC + 2];              // the all-zero record
  double rowc[16];                   // chunk limits and grid points of the row
  int    info[32];
  int    pbuf[32];
  int4   ent[MAXNS * RX + 1];        // merged stream, as in rowtile.cuh
};
static_assert(sizeof(WarpSmem) % 16 == 0, "the records of the next warp must stay 16-byte aligned");

constexpr size_t SMEM_BYTES = sizeof(double) * FTILE + sizeof(BlockSmem) + sizeof(WarpSmem) * WARPS;

template <int NY, int NX>
__device__ __forceinline__ double interp_cell(const double* __restrict__ p, const double* wy, const double* wx)
{
  double ry = 0;
#pragma unroll
  for (int jy = 0; jy < NY; jy++) {
    double rx = 0;
#pragma unroll
    for (int jx = 0; jx < NX; jx++)
      rx += p[jy * FROW + jx * 6] * wx[jx];
    ry += rx * wy[jy];
  }
  return ry;
}

// phase 1: stage the factors of one particle; qvz = q (z_new - z_old) / dt
__device__ __forceinline__ void stage_particle(double* __restrict__ rec, const AxisFactors& fx,
                                               const AxisFactors& fy, double q, double qvz, double dxdt,
                                               double dydt)
{
  const double A = 1.0 / 2, B = 1.0 / 3;
  const double cx = -q * dxdt, cy = -q * dydt;
  double       ax[4], wy[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    ax[k] = fx.S0[k] + A * fx.DS[k];
    wy[k] = fy.S0[k] + A * fy.DS[k];
  }
  store2(rec + T_P + 0, fy.S1[0], fy.S1[1]);
  store2(rec + T_P + 2, fy.S1[2], fy.S1[3]);
  store2(rec + T_P + 4, wy[0], wy[1]);
  store2(rec + T_P + 6, wy[2], wy[3]);
  store2(rec + T_P + 8, ax[0], ax[1]);
  store2(rec + T_P + 10, ax[2], ax[3]);
  store2(rec + T_P + 12, fy.S0[0], fy.S0[1]);
  store2(rec + T_P + 14, fy.S0[2], fy.S0[3]);
  store2(rec + T_Q + 0, fy.DS[0], fy.DS[1]);
  store2(rec + T_Q + 2, fy.DS[2], fy.DS[3]);
  const double px0 = fx.DS[0], px1 = px0 + fx.DS[1], px2 = px1 + fx.DS[2];
  const double py0 = fy.DS[0], py1 = py0 + fy.DS[1], py2 = py1 + fy.DS[2];
  store2(rec + T_Q + 4, 0.0, 0.0);
  store2(rec + T_R1 + 0, q * fx.S1[0], q * fx.S1[1]);
  store2(rec + T_R1 + 2, q * fx.S1[2], q * fx.S1[3]);
  store2(rec + T_R1 + 4, cx * px0, cx * px1);
  store2(rec + T_R1 + 6, cx * px2, 0.0);
  store2(rec + T_R1 + 8, cy * py0, cy * py1);
  store2(rec + T_R1 + 10, cy * py2, 0.0);
  store2(rec + T_R1 + 12, qvz * ax[0], qvz * ax[1]);
  store2(rec + T_R1 + 14, qvz * ax[2], qvz * ax[3]);
#pragma unroll
  for (int k = 0; k < 4; k += 2)
    store2(rec + T_R2 + k, qvz * (A * fx.S0[k] + B * fx.DS[k]), qvz * (A * fx.S0[k + 1] + B * fx.DS[k + 1]));
}

// lane constants of phase 2: lane (c, a) = (component, line) inside its half-warp
struct LaneMap {
  int p, q;  // record offsets of P and Q
  int r1, r2;
  int lin;   // lane part of the tile index (the run adds wy*SY + jx + wx)
  int sj;    // tile stride of j, in elements (x 4 components)
  int c;
};

__device__ __forceinline__ LaneMap lane_map(int lane)
{
  const int a = lane & 3;
  const int c = (lane >> 2) & 3;
  LaneMap   m;
  m.c   = c;
  m.p   = T_P + 4 * c + a;
  m.q   = c == 3 ? T_Q + a : T_Q + 4;
  m.r1  = T_R1 + 4 * c;
  m.r2  = c == 3 ? T_R2 : T_R1 + 4 * c;
  m.lin = c == 0 ? a * SY : (c == 1 ? a * SY + 1 : (c == 2 ? SY + a : a * SY));
  m.sj  = 4 * (c == 2 ? SY : 1);
  return m;
}

struct Acc {
  double v[4];
  __device__ __forceinline__ void clear() { v[0] = v[1] = v[2] = v[3] = 0; }
};

__device__ __forceinline__ void accumulate(Acc& acc, const double* __restrict__ rec, const LaneMap& m)
{
  const double  P = rec[m.p], Q = rec[m.q];
  const double2 ra = *reinterpret_cast<const double2*>(rec + m.r1);
  const double2 rb = *reinterpret_cast<const double2*>(rec + m.r1 + 2);
  const double2 sa = *reinterpret_cast<const double2*>(rec + m.r2);
  const double2 sb = *reinterpret_cast<const double2*>(rec + m.r2 + 2);
  acc.v[0] += P * ra.x + Q * sa.x;
  acc.v[1] += P * ra.y + Q * sa.y;
  acc.v[2] += P * rb.x + Q * sb.x;
  acc.v[3] += P * rb.y + Q * sb.y;
}

// both half-warps hold accumulators of the same cell: each keeps two of the four j and receives the
// partner's contribution to them; the current components with a prefix sum only have j < 3
__device__ __forceinline__ void flush(double* __restrict__ tile, const Acc& acc, const LaneMap& m, int run,
                                      int half)
{
  double* p = tile + 4 * (m.lin + run) + m.c + 2 * half * m.sj;
#pragma unroll
  for (int jj = 0; jj < 2; jj++) {
    const double give = half ? acc.v[jj] : acc.v[2 + jj];
    const double keep = half ? acc.v[2 + jj] : acc.v[jj];
    const double sum  = keep + __shfl_xor_sync(0xffffffffu, give, 16);
    const bool   real = m.c == 0 || m.c == 3 || 2 * half + jj < 3;
    if (real)
      p[jj * m.sj] += sum;
  }
}

// one staged particle straight into the tile, the whole warp on it (each half-warp two of the four j)
__device__ __forceinline__ void deposit_direct(double* __restrict__ tile, const double* __restrict__ rec,
                                               const LaneMap& m, int run, int half)
{
  const double P = rec[m.p], Q = rec[m.q];
  double*      p = tile + 4 * (m.lin + run) + m.c + 2 * half * m.sj;
#pragma unroll
  for (int jj = 0; jj < 2; jj++) {
    const int j = 2 * half + jj;
    if (m.c == 0 || m.c == 3 || j < 3)
      p[jj * m.sj] += P * rec[m.r1 + j] + Q * rec[m.r2 + j];
  }
}

__device__ __forceinline__ int run_index(int info)
{
  const int jx = info & 0xff;
  const int wx = (info >> 8) & 1, wy = (info >> 9) & 1;
  return wy * SY + jx + wx;
}

} // namespace rowtile2d
} // namespace picnix

#endif
