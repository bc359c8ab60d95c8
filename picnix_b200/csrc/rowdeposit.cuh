// -*- C++ -*-
// Row-owner Esirkepov deposit for 3-D, 2nd-order shapes (the BASELINE "T3D" configuration).
//
// Why: one fp64 atomic per stencil value (up to 5^3 x 4 per particle) is two orders of magnitude
// too slow -- shared-memory fp64 atomics are CAS loops on sm_100a and L2 reductions saturate near
// 1e11/s.  So nothing in the inner loop is atomic.  Instead
//
//   * a warp owns a ROW segment of RX cells (fixed z, y) and walks its cell-sorted particles;
//   * phase 1 (thread per particle): push (optional), then the 1-D Esirkepov factors of the
//     particle on a 4-slot WINDOW per axis -- for a 2nd-order shape and |move| < 1 cell the old and
//     new weights together never span more than 4 of the 5 stencil slots
//     (nix/esirkepov.hpp:240-275: the new weights are the old stencil shifted by -1/0/+1) --
//     written to a per-warp staging buffer in shared memory;
//   * phase 2 (thread per stencil point): each half-warp takes one staged particle; lane (a,b)
//     owns the window points (a,b,*) and accumulates rho/Jx/Jy/Jz in REGISTERS with one FMA per
//     value (the value is never materialised), exactly the sums of nix/esirkepov.hpp:154-237:
//         rho[z][y][x]   += (q S1z[z] S1y[y])        * S1x[x]
//         Jx [z][y][x+1] += W(y,z) * prefix_x DSx,   W = -q dx/dt ((S0y+DSy/2) S0z + (S0y/2+DSy/3) DSz)
//         (Jy, Jz by cyclic permutation)
//   * when the cell changes the 13 registers of a lane are added -- plain loads/stores, the lanes
//     own distinct points -- into the warp's PRIVATE (RX+4) x 5 x 5 x 4 tile in shared memory;
//   * at the end of the row segment the tile goes to global uj with one fp64 reduction per tile
//     value (about 2.3 per particle at 64 ppc instead of ~170).
//
// The few particles whose new weights fall left of the window (cell shift -1) use the same code
// with a window that starts one slot lower and are flushed individually.  Results differ from the
// reference only by summation order (tolerance in tests: 1e-12 of max |J|).
#ifndef PICNIX_B200_ROWDEPOSIT_CUH
#define PICNIX_B200_ROWDEPOSIT_CUH

#include "particle_kernels.cuh"

namespace picnix
{
namespace rowdep
{

constexpr int RX       = 8;               // cells per row segment
constexpr int XS       = RX + 4;          // tile extent in x (stencil reaches -2..+2)
constexpr int SY       = XS * 4 + 1;      // tile stride of y in doubles   (== 1 mod 16)
constexpr int SZ       = 5 * SY + 15;     // tile stride of z in doubles   (== 4 mod 16)
constexpr int TILE     = 5 * SZ;          // doubles per warp tile
constexpr int NSTG     = 53;              // staged doubles per particle (odd: conflict-poor)
constexpr int WARPS    = 4;               // warps per block
constexpr int THREADS  = WARPS * 32;

// offsets inside a staged particle record
constexpr int O_S1X = 0;   // S1x[4]            new x weights on the window
constexpr int O_PX  = 4;   // Px[3]             -q dx/dt * prefix sums of DSx (window idx 1..3)
constexpr int O_PY  = 7;   // Py[3]
constexpr int O_PZ  = 10;  // Pz[3]
constexpr int O_QS1Z = 13; // q * S1z[4]
constexpr int O_S0Z = 17;  // S0z[4]
constexpr int O_DSZ = 21;  // DSz[4]
constexpr int O_S0Y = 25;  // S0y[4]
constexpr int O_DSY = 29;  // DSy[4]
constexpr int O_S1Y = 33;  // S1y[4]
constexpr int O_AY  = 37;  // S0y + DSy/2
constexpr int O_BY  = 41;  // S0y/2 + DSy/3
constexpr int O_AX  = 45;  // S0x + DSx/2
constexpr int O_BX  = 49;  // S0x/2 + DSx/3

struct WarpSmem {
  double stg[32 * NSTG];
  double tile[TILE];
  int    info[32];
};

constexpr size_t SMEM_BYTES = sizeof(WarpSmem) * WARPS;

// info word: bits 0..7 cell index inside the segment, bit 8/9/10 window offset x/y/z (1 = majority
// window that starts at the old cell's slot 1), bit 11 valid
__device__ __forceinline__ int make_info(int jx, int wx, int wy, int wz)
{
  return jx | (wx << 8) | (wy << 9) | (wz << 10) | (1 << 11);
}

// One axis: old/new 2nd-order weights on the 4-slot window.  Returns false when the move cannot be
// represented (shift beyond one cell), in which case the caller takes the generic path.
struct AxisFactors {
  double S0[4], S1[4], DS[4];
  int    i0;  // old cell (relative to the chunk)
  int    w;   // window offset: 1 = slots 1..4, 0 = slots 0..3
  bool   ok;
};

__device__ __forceinline__ AxisFactors axis_factors(double x0, double x1, double xmin, double dx)
{
  AxisFactors  f;
  const double rdx   = 1 / dx;
  const double xgrid = xmin + 0.5 * dx;

  f.i0 = digitize(x0, xmin, rdx);
  double s0[3], s1[3];
  shape_mc<2>(x0, xgrid + (double)f.i0 * dx, rdx, s0);
  const int i1 = digitize(x1, xmin, rdx);
  shape_mc<2>(x1, xgrid + (double)i1 * dx, rdx, s1);

  const int sh = i1 - f.i0;
  f.ok         = sh >= -1 && sh <= 1;
  f.w          = sh < 0 ? 0 : 1;
  const bool w1 = f.w == 1;
  const bool up = sh > 0; // new weights one slot to the right inside the window

  f.S0[0] = w1 ? s0[0] : 0.0;
  f.S0[1] = w1 ? s0[1] : s0[0];
  f.S0[2] = w1 ? s0[2] : s0[1];
  f.S0[3] = w1 ? 0.0 : s0[2];
  f.S1[0] = up ? 0.0 : s1[0];
  f.S1[1] = up ? s1[0] : s1[1];
  f.S1[2] = up ? s1[1] : s1[2];
  f.S1[3] = up ? s1[2] : 0.0;
#pragma unroll
  for (int k = 0; k < 4; k++)
    f.DS[k] = f.S1[k] - f.S0[k];
  return f;
}

// phase 1: stage the factors of one particle (lane-private record)
__device__ __forceinline__ void stage_particle(double* __restrict__ rec, const AxisFactors& fx,
                                               const AxisFactors& fy, const AxisFactors& fz,
                                               double q, double dxdt, double dydt, double dzdt)
{
  const double A = 1.0 / 2, B = 1.0 / 3;
  const double cx = -q * dxdt, cy = -q * dydt, cz = -q * dzdt;
  double       p;
#pragma unroll
  for (int k = 0; k < 4; k++)
    rec[O_S1X + k] = fx.S1[k];
  p = fx.DS[0];
  rec[O_PX + 0] = cx * p;
  p += fx.DS[1];
  rec[O_PX + 1] = cx * p;
  p += fx.DS[2];
  rec[O_PX + 2] = cx * p;
  p = fy.DS[0];
  rec[O_PY + 0] = cy * p;
  p += fy.DS[1];
  rec[O_PY + 1] = cy * p;
  p += fy.DS[2];
  rec[O_PY + 2] = cy * p;
  p = fz.DS[0];
  rec[O_PZ + 0] = cz * p;
  p += fz.DS[1];
  rec[O_PZ + 1] = cz * p;
  p += fz.DS[2];
  rec[O_PZ + 2] = cz * p;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    rec[O_QS1Z + k] = q * fz.S1[k];
    rec[O_S0Z + k]  = fz.S0[k];
    rec[O_DSZ + k]  = fz.DS[k];
    rec[O_S0Y + k]  = fy.S0[k];
    rec[O_DSY + k]  = fy.DS[k];
    rec[O_S1Y + k]  = fy.S1[k];
    rec[O_AY + k]   = fy.S0[k] + A * fy.DS[k];
    rec[O_BY + k]   = A * fy.S0[k] + B * fy.DS[k];
    rec[O_AX + k]   = fx.S0[k] + A * fx.DS[k];
    rec[O_BX + k]   = A * fx.S0[k] + B * fx.DS[k];
  }
}

// the 13 register accumulators of a lane
struct Acc {
  double rho[4], jx[3], jy[3], jz[3];
  __device__ __forceinline__ void clear()
  {
#pragma unroll
    for (int k = 0; k < 4; k++)
      rho[k] = 0;
#pragma unroll
    for (int k = 0; k < 3; k++)
      jx[k] = jy[k] = jz[k] = 0;
  }
};

// phase 2 body: contributions of one staged particle to the points owned by lane (a, b)
__device__ __forceinline__ void accumulate(Acc& acc, const double* __restrict__ rec, int a, int b)
{
  const double c   = rec[O_QS1Z + a] * rec[O_S1Y + b];
  const double s0z = rec[O_S0Z + a], dsz = rec[O_DSZ + a];
  const double s0y = rec[O_S0Y + a], dsy = rec[O_DSY + a];
  const double wyz = rec[O_AY + b] * s0z + rec[O_BY + b] * dsz; // (jz=a, jy=b)
  const double wzx = rec[O_AX + b] * s0z + rec[O_BX + b] * dsz; // (jz=a, jx=b)
  const double wxy = rec[O_AX + b] * s0y + rec[O_BX + b] * dsy; // (jy=a, jx=b)
#pragma unroll
  for (int k = 0; k < 4; k++)
    acc.rho[k] += c * rec[O_S1X + k];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    acc.jx[k] += wyz * rec[O_PX + k];
    acc.jy[k] += wzx * rec[O_PY + k];
    acc.jz[k] += wxy * rec[O_PZ + k];
  }
}

// add a lane's accumulators into the warp tile; (wz,wy,wx) window offset, jx cell in the segment
__device__ __forceinline__ void flush(double* __restrict__ tile, const Acc& acc, int a, int b,
                                      int jx, int wx, int wy, int wz)
{
  // rho and Jx: point (z = wz+a, y = wy+b, x = jx+wx+k)
  double* p = tile + (wz + a) * SZ + (wy + b) * SY + (jx + wx) * 4;
#pragma unroll
  for (int k = 0; k < 4; k++)
    p[k * 4 + 0] += acc.rho[k];
#pragma unroll
  for (int k = 0; k < 3; k++)
    p[(k + 1) * 4 + 1] += acc.jx[k];
  // Jy: point (z = wz+a, y = wy+k+1, x = jx+wx+b)
  double* py = tile + (wz + a) * SZ + wy * SY + (jx + wx + b) * 4 + 2;
#pragma unroll
  for (int k = 0; k < 3; k++)
    py[(k + 1) * SY] += acc.jy[k];
  // Jz: point (z = wz+k+1, y = wy+a, x = jx+wx+b)
  double* pz = tile + wz * SZ + (wy + a) * SY + (jx + wx + b) * 4 + 3;
#pragma unroll
  for (int k = 0; k < 3; k++)
    pz[(k + 1) * SZ] += acc.jz[k];
}

} // namespace rowdep
} // namespace picnix

#endif
