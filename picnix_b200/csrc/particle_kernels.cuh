// -*- C++ -*-
// Per-particle device routines: field interpolation + momentum push (K1) and Esirkepov
// charge-conserving current deposition (K2), templated on dimensionality and shape order.
//
// Reference:
//   BaseVelocity::weights{1,2,3}d / push_scalar{1,2,3}d   pic/engine/velocity.hpp:222-450
//   interp{1,2,3}d                                        nix/interp.hpp:14-113
//   BaseCurrent::local{1,2,3}d                            pic/engine/current.hpp:207-395
//   esirkepov::{shift_weights,deposit{1,2,3}d}            nix/esirkepov.hpp:18-340
#ifndef PICNIX_B200_PARTICLE_KERNELS_CUH
#define PICNIX_B200_PARTICLE_KERNELS_CUH

#include "particle_common.cuh"

namespace picnix
{

// Weights of one axis for the velocity push: integer-grid (cell-centred quantities) and
// half-grid (edge quantities) weights and first stencil indices (already shifted to array
// indices: + lb - Order/2).
template <int Order, int Interp>
__device__ __forceinline__ void axis_weights(double x, double xmin, double dx, double cfl, int lb,
                                             double* wi, double* wh, int& i0, int& h0)
{
  constexpr int is_odd = Order % 2;
  const double  ximin  = xmin + 0.5 * dx * is_odd;
  const double  xhmin  = xmin + 0.5 * dx * is_odd - 0.5 * dx;
  const double  xigrid = xmin + 0.5 * dx;
  const double  xhgrid = xmin;
  const double  rdx    = 1 / dx;

  i0 = digitize(x, ximin, rdx);
  h0 = digitize(x, xhmin, rdx);
  const double xig = xigrid + (double)i0 * dx;
  const double xhg = xhgrid + (double)h0 * dx;

  if (Interp == PICNIX_INTERP_MC) {
    shape_mc<Order>(x, xig, rdx, wi);
  } else {
    shape_wt<Order>(x, xig, rdx, cfl, 1 / cfl, wi);
  }
  shape_mc<Order>(x, xhg, rdx, wh);

  i0 += lb - (Order / 2);
  h0 += lb - (Order / 2);
}

// tensor-product interpolation of component k, x innermost (nix/interp.hpp:94-113);
// F is any accessor returning the field value at (iz, iy, ix, k)
template <int Dim, int Order, typename Field>
__device__ __forceinline__ double interpolate(const Field& F, int iz0, int iy0, int ix0, int k,
                                              const double* wz, const double* wy,
                                              const double* wx, double dt)
{
  constexpr int N = Order + 1;
  if (Dim == 1) {
    double rx = 0;
#pragma unroll
    for (int jx = 0; jx < N; jx++)
      rx += F(iz0, iy0, ix0 + jx, k) * wx[jx];
    return rx * dt;
  } else if (Dim == 2) {
    double ry = 0;
#pragma unroll
    for (int jy = 0; jy < N; jy++) {
      double rx = 0;
#pragma unroll
      for (int jx = 0; jx < N; jx++)
        rx += F(iz0, iy0 + jy, ix0 + jx, k) * wx[jx];
      ry += rx * wy[jy];
    }
    return ry * dt;
  } else {
    double rz = 0;
#pragma unroll
    for (int jz = 0; jz < N; jz++) {
      double ry = 0;
#pragma unroll
      for (int jy = 0; jy < N; jy++) {
        double rx = 0;
#pragma unroll
        for (int jx = 0; jx < N; jx++)
          rx += F(iz0 + jz, iy0 + jy, ix0 + jx, k) * wx[jx];
        ry += rx * wy[jy];
      }
      rz += ry * wz[jz];
    }
    return rz * dt;
  }
}

// E and B at the particle, pre-multiplied by qmdt, then the momentum update.
// lim = chunk limits {zmin,zmax,ymin,ymax,xmin,xmax}.
template <int Dim, int Order, int Pusher, int Interp, typename Field>
__device__ __forceinline__ void velocity_update(const Geom& g, const double* __restrict__ lim,
                                                const Field& F, double delt, double qmdt,
                                                double x, double y, double z, double& ux,
                                                double& uy, double& uz)
{
  constexpr int N = Order + 1;
  double wix[N], whx[N], wiy[N], why[N], wiz[N], whz[N];
  int    ix0, hx0, iy0 = g.Lb[1], hy0 = g.Lb[1], iz0 = g.Lb[0], hz0 = g.Lb[0];

  axis_weights<Order, Interp>(x, lim[4], g.del[2], g.cc * delt / g.del[2], g.Lb[2], wix, whx, ix0,
                              hx0);
  if (Dim >= 2)
    axis_weights<Order, Interp>(y, lim[2], g.del[1], g.cc * delt / g.del[1], g.Lb[1], wiy, why,
                                iy0, hy0);
  if (Dim >= 3)
    axis_weights<Order, Interp>(z, lim[0], g.del[0], g.cc * delt / g.del[0], g.Lb[0], wiz, whz,
                                iz0, hz0);

  // Yee staggering, pic/engine/velocity.hpp:379-384, 410-415, 442-447
  double ex = interpolate<Dim, Order>(F, iz0, iy0, hx0, 0, wiz, wiy, whx, qmdt);
  double ey = interpolate<Dim, Order>(F, iz0, hy0, ix0, 1, wiz, why, wix, qmdt);
  double ez = interpolate<Dim, Order>(F, hz0, iy0, ix0, 2, whz, wiy, wix, qmdt);
  double bx = interpolate<Dim, Order>(F, hz0, hy0, ix0, 3, whz, why, wix, qmdt);
  double by = interpolate<Dim, Order>(F, hz0, iy0, hx0, 4, whz, wiy, whx, qmdt);
  double bz = interpolate<Dim, Order>(F, iz0, hy0, hx0, 5, wiz, why, whx, qmdt);

  push_momentum<Pusher>(ux, uy, uz, ex, ey, ez, bx, by, bz, g.cc);
}

//
// Esirkepov density decomposition.
//
// ss[0][d][*]: weights before the move, ss[1][d][*]: after the move (both on Order+3 slots, the
// Order+1 weights written from slot 1, the "after" set shifted by the cell displacement).
// Returns the first array index of the stencil in each direction.
//
template <int Order>
__device__ __forceinline__ int esirkepov_axis(double x0, double x1, double xmin, double dx, int lb,
                                              double* s0, double* s1)
{
  constexpr int S      = Order + 3;
  constexpr int is_odd = Order % 2;
  const double  rdx    = 1 / dx;
  const double  ximin  = xmin + 0.5 * dx * is_odd;
  const double  xgrid  = xmin + 0.5 * dx;

#pragma unroll
  for (int j = 0; j < S; j++) {
    s0[j] = 0;
    s1[j] = 0;
  }

  const int    i0  = digitize(x0, ximin, rdx);
  const double xg0 = xgrid + (double)i0 * dx;
  shape_mc<Order>(x0, xg0, rdx, s0 + 1);

  const int    i1  = digitize(x1, ximin, rdx);
  const double xg1 = xgrid + (double)i1 * dx;
  shape_mc<Order>(x1, xg1, rdx, s1 + 1);

  // esirkepov::shift_weights, nix/esirkepov.hpp:240-259
  const int shift = i1 - i0;
  if (shift < 0) {
#pragma unroll
    for (int j = 0; j < S - 1; j++)
      s1[j] = s1[j + 1];
  } else if (shift > 0) {
#pragma unroll
    for (int j = S - 1; j > 0; j--)
      s1[j] = s1[j - 1];
  }

  return i0 + lb - (Order / 2) - 1;
}

// Deposit one particle through `add(jz, jy, jx, k, value)` where (jz,jy,jx) are stencil offsets.
// The operation order inside each running sum follows nix/esirkepov.hpp exactly.
template <int Dim, int Order, typename Add>
__device__ __forceinline__ void esirkepov_deposit(const Geom& g, const double* __restrict__ lim,
                                                  double q, double delt, double x0, double y0,
                                                  double z0, double x1, double y1, double z1,
                                                  int& bz, int& by, int& bx, const Add& add)
{
  constexpr int S = Order + 3;
  const double  A = 1.0 / 2, B = 1.0 / 3;

  double sx0[S], sx1[S];
  bx = esirkepov_axis<Order>(x0, x1, lim[4], g.del[2], g.Lb[2], sx0, sx1);
  by = g.Lb[1];
  bz = g.Lb[0];

  if (Dim == 1) {
    const double vy = (y1 - y0) / delt;
    const double vz = (z1 - z0) / delt;
    // rho from the NEW weights, then DS = S1 - S0
#pragma unroll
    for (int jx = 0; jx < S; jx++)
      add(0, 0, jx, 0, q * sx1[jx]);
#pragma unroll
    for (int jx = 0; jx < S; jx++)
      sx1[jx] -= sx0[jx];

    const double qdxdt = q * (g.del[2] / delt);
    double       ww    = 0;
    const double wx    = -qdxdt;
#pragma unroll
    for (int jx = 0; jx < S - 1; jx++) {
      ww += sx1[jx] * wx;
      add(0, 0, jx + 1, 1, ww);
    }
    const double qvy = q * vy, qvz = q * vz;
#pragma unroll
    for (int jx = 0; jx < S; jx++) {
      add(0, 0, jx, 2, (sx0[jx] + A * sx1[jx]) * qvy);
      add(0, 0, jx, 3, (sx0[jx] + A * sx1[jx]) * qvz);
    }
    return;
  }

  double sy0[S], sy1[S];
  by = esirkepov_axis<Order>(y0, y1, lim[2], g.del[1], g.Lb[1], sy0, sy1);

  if (Dim == 2) {
    const double vz = (z1 - z0) / delt;
#pragma unroll
    for (int jy = 0; jy < S; jy++)
#pragma unroll
      for (int jx = 0; jx < S; jx++)
        add(0, jy, jx, 0, q * sx1[jx] * sy1[jy]);
#pragma unroll
    for (int j = 0; j < S; j++) {
      sx1[j] -= sx0[j];
      sy1[j] -= sy0[j];
    }

    const double qdxdt = q * (g.del[2] / delt);
    const double qdydt = q * (g.del[1] / delt);
    const double qvz   = q * vz;
#pragma unroll
    for (int jy = 0; jy < S; jy++) {
      double       ww = 0;
      const double wx = -(sy0[jy] + A * sy1[jy]) * qdxdt;
#pragma unroll
      for (int jx = 0; jx < S - 1; jx++) {
        ww += sx1[jx] * wx;
        add(0, jy, jx + 1, 1, ww);
      }
    }
#pragma unroll
    for (int jx = 0; jx < S; jx++) {
      double       ww = 0;
      const double wy = -(sx0[jx] + A * sx1[jx]) * qdydt;
#pragma unroll
      for (int jy = 0; jy < S - 1; jy++) {
        ww += sy1[jy] * wy;
        add(0, jy + 1, jx, 2, ww);
      }
    }
#pragma unroll
    for (int jy = 0; jy < S; jy++)
#pragma unroll
      for (int jx = 0; jx < S; jx++)
        add(0, jy, jx, 3,
            ((1 * sx0[jx] + A * sx1[jx]) * sy0[jy] + (A * sx0[jx] + B * sx1[jx]) * sy1[jy]) * qvz);
    return;
  }

  double sz0[S], sz1[S];
  bz = esirkepov_axis<Order>(z0, z1, lim[0], g.del[0], g.Lb[0], sz0, sz1);

#pragma unroll
  for (int jz = 0; jz < S; jz++)
#pragma unroll
    for (int jy = 0; jy < S; jy++)
#pragma unroll
      for (int jx = 0; jx < S; jx++)
        add(jz, jy, jx, 0, q * sx1[jx] * sy1[jy] * sz1[jz]);

#pragma unroll
  for (int j = 0; j < S; j++) {
    sx1[j] -= sx0[j];
    sy1[j] -= sy0[j];
    sz1[j] -= sz0[j];
  }

  const double qdxdt = q * (g.del[2] / delt);
  const double qdydt = q * (g.del[1] / delt);
  const double qdzdt = q * (g.del[0] / delt);

#pragma unroll
  for (int jz = 0; jz < S; jz++)
#pragma unroll
    for (int jy = 0; jy < S; jy++) {
      double       ww = 0;
      const double wx = -((1 * sy0[jy] + A * sy1[jy]) * sz0[jz] +
                          (A * sy0[jy] + B * sy1[jy]) * sz1[jz]) * qdxdt;
#pragma unroll
      for (int jx = 0; jx < S - 1; jx++) {
        ww += sx1[jx] * wx;
        add(jz, jy, jx + 1, 1, ww);
      }
    }
#pragma unroll
  for (int jz = 0; jz < S; jz++)
#pragma unroll
    for (int jx = 0; jx < S; jx++) {
      double       ww = 0;
      const double wy = -((1 * sz0[jz] + A * sz1[jz]) * sx0[jx] +
                          (A * sz0[jz] + B * sz1[jz]) * sx1[jx]) * qdydt;
#pragma unroll
      for (int jy = 0; jy < S - 1; jy++) {
        ww += sy1[jy] * wy;
        add(jz, jy + 1, jx, 2, ww);
      }
    }
#pragma unroll
  for (int jy = 0; jy < S; jy++)
#pragma unroll
    for (int jx = 0; jx < S; jx++) {
      double       ww = 0;
      const double wz = -((1 * sx0[jx] + A * sx1[jx]) * sy0[jy] +
                          (A * sx0[jx] + B * sx1[jx]) * sy1[jy]) * qdzdt;
#pragma unroll
      for (int jz = 0; jz < S - 1; jz++) {
        ww += sz1[jz] * wz;
        add(jz + 1, jy, jx, 3, ww);
      }
    }
}

} // namespace picnix

#endif
