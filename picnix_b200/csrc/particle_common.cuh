// -*- C++ -*-
// Device building blocks of the particle kernels: cell lookup, shape functions, momentum pushers
// and the per-dimension weight set-up shared by the velocity push and the current deposit.
//
// The pushers and the WT shape functions below are closed-form formulas whose operation order is pinned
// by the parity bar (1e-13 against the reference on identical inputs): they follow nix/primitives.hpp
// term by term, including its temporaries, because any re-association shows up in the last digits.
//
// Reference (all scalar forms; the xsimd forms compute the same numbers lane-wise):
//   digitize                     nix/primitives.hpp:45-58
//   lorentz_factor               nix/primitives.hpp:157-161
//   push_boris/vay/higuera_cary  nix/primitives.hpp:163-253
//   shape_mc<1..4>               nix/primitives.hpp:255-329
//   shape_wt<1..4>               nix/primitives.hpp:331-495
//   cell key (count)             nix/xtensor_particle.hpp:324-357
#ifndef PICNIX_B200_PARTICLE_COMMON_CUH
#define PICNIX_B200_PARTICLE_COMMON_CUH

#include "arena.hpp"

#include <cfloat>

namespace picnix
{

// remember a particle that left its chunk (key == Ng) for the migration kernel
__device__ __forceinline__ void note_leaver(const DevPtrs& d, int seg, int64_t slot)
{
  const int k = atomicAdd(d.leave_count, 1);
  if (k < d.leave_cap)
    d.leave_idx[k] = ((int64_t)seg << 40) | slot;
}

__device__ __forceinline__ int digitize(double x, double xmin, double rdx)
{
  return (int)floor((x - xmin) * rdx);
}

//
// Momentum-conserving shape functions: weights of the Order+1 grid points around X
//
template <int Order>
__device__ __forceinline__ void shape_mc(double x, double X, double rdx, double* s)
{
  const double delta = (x - X) * rdx;
  if (Order == 1) {
    s[0] = 1 - delta;
    s[1] = delta;
  } else if (Order == 2) {
    const double w1 = 0.5 - delta;
    const double w2 = 0.5 + delta;
    s[0] = 0.50 * w1 * w1;
    s[1] = 0.75 - delta * delta;
    s[2] = 0.50 * w2 * w2;
  } else if (Order == 3) {
    const double a  = 1 / 6.0;
    const double w1 = delta;
    const double w2 = 1 - delta;
    const double p1 = w1 * w1, p2 = w2 * w2;
    const double c1 = p1 * w1, c2 = p2 * w2;
    s[0] = a * c2;
    s[1] = a * (4 - 6 * p1 + 3 * c1);
    s[2] = a * (4 - 6 * p2 + 3 * c2);
    s[3] = a * c1;
  } else {
    const double a = 1 / 384.0, b = 1 / 96.0, c = 115 / 192.0, d = 1 / 8.0;
    const double w1 = 1 + delta, w2 = 1 - delta;
    const double w3 = 1 + delta * 2, w4 = 1 - delta * 2;
    const double q0 = delta * delta;
    const double q1 = w1 * w1, q2 = w2 * w2;
    const double c1 = q1 * w1, c2 = q2 * w2;
    const double f1 = c1 * w1, f2 = c2 * w2;
    const double f3 = w3 * w3 * w3 * w3, f4 = w4 * w4 * w4 * w4;
    s[0] = a * f4;
    s[1] = b * (55 + 20 * w1 - 120 * q1 + 80 * c1 - 16 * f1);
    s[2] = c + d * q0 * (2 * q0 - 5);
    s[3] = b * (55 + 20 * w2 - 120 * q2 + 80 * c2 - 16 * f2);
    s[4] = a * f3;
  }
}

//
// WT-scheme shape functions (Lu et al. 2020): time-step dependent weights, dt = c*delt/dx
//
template <int Order>
__device__ __forceinline__ void shape_wt(double x, double X, double rdx, double dt, double rdt,
                                         double* s)
{
  const double delta = (x - X) * rdx;
  if (Order == 1) {
    const double ss = fmin(1.0, fmax(0.0, 0.25 * rdt * (1 + 2 * dt - 2 * delta)));
    s[0] = ss;
    s[1] = 1 - ss;
  } else if (Order == 2) {
    const double t1 = delta < -dt ? 1.0 : 0.0;
    const double t2 = 1 - t1;
    const double t3 = delta < +dt ? 1.0 : 0.0;
    const double t4 = 1 - t3;
    const double w0 = fabs(delta);
    const double w1 = dt - delta;
    const double w2 = dt + delta;

    const double s0_1 = w0, s1_1 = 1 - w0, s2_1 = 0;
    const double s0_2 = 0.25 * rdt * w1 * w1;
    const double s1_2 = 0.50 * rdt * (dt * (2 - dt) - w0 * w0);
    const double s2_2 = 0.25 * rdt * w2 * w2;
    const double s0_3 = s2_1, s1_3 = s1_1, s2_3 = s0_1;

    s[0] = s0_1 * t1 + s0_2 * t2 * t3 + s0_3 * t4;
    s[1] = s1_1 * t1 + s1_2 * t2 * t3 + s1_3 * t4;
    s[2] = s2_1 * t1 + s2_2 * t2 * t3 + s2_3 * t4;
  } else if (Order == 3) {
    const double a = 1 / 96.0, b = 1 / 24.0, c = 1 / 12.0;
    const double adt = a * rdt;

    const double t1 = delta < 0.5 - dt ? 1.0 : 0.0;
    const double t2 = 1 - t1;
    const double t3 = delta < 0.5 + dt ? 1.0 : 0.0;
    const double t4 = 1 - t3;
    const double w0 = delta;
    const double w1 = 1 - delta;
    const double w3 = 1 - 2 * delta;
    const double w4 = 1 + 2 * delta;
    const double w5 = 2 * dt + w3;
    const double w6 = 2 * dt - w3;
    const double w7 = 3 - 2 * delta;
    const double w0_2 = w0 * w0, w1_2 = w1 * w1;
    const double w3_2 = w3 * w3, w3_3 = w3_2 * w3;
    const double w4_2 = w4 * w4;
    const double w5_3 = w5 * w5 * w5, w6_3 = w6 * w6 * w6;
    const double w7_2 = w7 * w7;
    const double dt_2 = dt * dt, dt_3 = dt_2 * dt;
    const double dt_2_4 = 4 * dt_2;
    const double s_2_odd  = adt * (-8 * dt_3 - 6 * dt * w3_2);
    const double s_2_even = adt * (-36 * dt_2 * w3 - 3 * w3_3);

    const double s0_1 = b * (dt_2_4 + 3 * w3_2);
    const double s1_1 = c * (9 - dt_2_4 - 12 * w0_2);
    const double s2_1 = b * (dt_2_4 + 3 * w4_2);
    const double s3_1 = 0;
    const double s0_2 = adt * w5_3;
    const double s1_2 = s_2_odd + s_2_even + w1;
    const double s2_2 = s_2_odd - s_2_even + w0;
    const double s3_2 = adt * w6_3;
    const double s0_3 = 0;
    const double s1_3 = b * (dt_2_4 + 3 * w7_2);
    const double s2_3 = c * (9 - dt_2_4 - 12 * w1_2);
    const double s3_3 = b * (dt_2_4 + 3 * w3_2);

    s[0] = s0_1 * t1 + s0_2 * t2 * t3 + s0_3 * t4;
    s[1] = s1_1 * t1 + s1_2 * t2 * t3 + s1_3 * t4;
    s[2] = s2_1 * t1 + s2_2 * t2 * t3 + s2_3 * t4;
    s[3] = s3_1 * t1 + s3_2 * t2 * t3 + s3_3 * t4;
  } else {
    const double a = 1 / 48.0, b = 1 / 24.0, c = 1 / 12.0, d = 1 / 6.0;
    const double adt = a * rdt, bdt = b * rdt, cdt = c * rdt;

    const double t1 = delta < -dt ? 1.0 : 0.0;
    const double t2 = 1 - t1;
    const double t3 = delta < +dt ? 1.0 : 0.0;
    const double t4 = 1 - t3;
    const double w0 = fabs(delta);
    const double w1 = 1 - w0;
    const double w2 = 1 - delta;
    const double w3 = 1 + delta;
    const double w4 = dt - delta;
    const double w5 = dt + delta;
    const double w0_2 = w0 * w0, w0_3 = w0_2 * w0, w0_4 = w0_3 * w0;
    const double w1_2 = w1 * w1, w1_3 = w1_2 * w1;
    const double w2_3 = w2 * w2 * w2, w3_3 = w3 * w3 * w3;
    const double w4_4 = w4 * w4 * w4 * w4, w5_4 = w5 * w5 * w5 * w5;
    const double dt_2 = dt * dt, dt_3 = dt_2 * dt, dt_4 = dt_3 * dt;
    const double ss1 = -dt_4 - 6 * w0_2 * dt_2 - w0_4;
    const double ss2 =
        3 * dt_4 - 8 * dt_3 + 18 * w0_2 * dt_2 + (16 - 24 * w0_2) * dt + 3 * w0_4;

    const double s0_1 = d * w0 * (w0_2 + dt_2);
    const double s1_1 = d * (4 - 6 * w1_2 + 3 * w1_3 + (1 - 3 * w0) * dt_2);
    const double s2_1 = d * (4 - 6 * w0_2 + 3 * w0_3 - (2 - 3 * w0) * dt_2);
    const double s3_1 = d * w1 * (w1_2 + dt_2);
    const double s4_1 = 0;
    const double s0_2 = adt * w4_4;
    const double s1_2 = cdt * (ss1 + 2 * dt_3 * w3 + 2 * dt * (-6 * delta + w3_3));
    const double s2_2 = bdt * ss2;
    const double s3_2 = cdt * (ss1 + 2 * dt_3 * w2 + 2 * dt * (+6 * delta + w2_3));
    const double s4_2 = adt * w5_4;
    const double s0_3 = s4_1, s1_3 = s3_1, s2_3 = s2_1, s3_3 = s1_1, s4_3 = s0_1;

    s[0] = s0_1 * t1 + s0_2 * t2 * t3 + s0_3 * t4;
    s[1] = s1_1 * t1 + s1_2 * t2 * t3 + s1_3 * t4;
    s[2] = s2_1 * t1 + s2_2 * t2 * t3 + s2_3 * t4;
    s[3] = s3_1 * t1 + s3_2 * t2 * t3 + s3_3 * t4;
    s[4] = s4_1 * t1 + s4_2 * t2 * t3 + s4_3 * t4;
  }
}

//
// momentum pushers; e* and b* are already multiplied by (q/m) dt / 2
//
__device__ __forceinline__ void push_boris(double& ux, double& uy, double& uz, double ex, double ey,
                                           double ez, double bx, double by, double bz, double cc)
{
  ux += ex;
  uy += ey;
  uz += ez;

  const double gm = 1 / sqrt(cc * cc + ux * ux + uy * uy + uz * uz);
  bx *= gm;
  by *= gm;
  bz *= gm;
  const double bb = 2.0 / (1.0 + bx * bx + by * by + bz * bz);

  const double vx = ux + (uy * bz - uz * by);
  const double vy = uy + (uz * bx - ux * bz);
  const double vz = uz + (ux * by - uy * bx);

  ux += (vy * bz - vz * by) * bb + ex;
  uy += (vz * bx - vx * bz) * bb + ey;
  uz += (vx * by - vy * bx) * bb + ez;
}

__device__ __forceinline__ void push_vay(double& ux, double& uy, double& uz, double ex, double ey,
                                         double ez, double bx, double by, double bz, double cc)
{
  double gm = 1 / sqrt(cc * cc + ux * ux + uy * uy + uz * uz);
  const double vx = ux + 2 * ex + gm * (uy * bz - uz * by);
  const double vy = uy + 2 * ey + gm * (uz * bx - ux * bz);
  const double vz = uz + 2 * ez + gm * (ux * by - uy * bx);

  gm        = (cc * cc + vx * vx + vy * vy + vz * vz);
  double bb = bx * bx + by * by + bz * bz;
  double bu = bx * vx + by * vy + bz * vz;
  const double xx = gm - bb;
  const double yy = bb + bu * bu;
  gm = 1 / sqrt(0.5 * (xx + sqrt(xx * xx + 4 * yy)));

  bx *= gm;
  by *= gm;
  bz *= gm;
  bu = bx * vx + by * vy + bz * vz;
  bb = 1.0 / (1.0 + bx * bx + by * by + bz * bz);

  ux = (vx + bu * bx + (vy * bz - vz * by)) * bb;
  uy = (vy + bu * by + (vz * bx - vx * bz)) * bb;
  uz = (vz + bu * bz + (vx * by - vy * bx)) * bb;
}

__device__ __forceinline__ void push_higuera_cary(double& ux, double& uy, double& uz, double ex,
                                                  double ey, double ez, double bx, double by,
                                                  double bz, double cc)
{
  ux += ex;
  uy += ey;
  uz += ez;

  double gm = cc * cc + ux * ux + uy * uy + uz * uz;
  double bb = bx * bx + by * by + bz * bz;
  const double bu = bx * ux + by * uy + bz * uz;
  const double xx = gm - bb;
  const double yy = bb + bu * bu;
  gm = 1 / sqrt(0.5 * (xx + sqrt(xx * xx + 4 * yy)));

  bx *= gm;
  by *= gm;
  bz *= gm;
  bb = 2.0 / (1.0 + bx * bx + by * by + bz * bz);

  const double vx = ux + (uy * bz - uz * by);
  const double vy = uy + (uz * bx - ux * bz);
  const double vz = uz + (ux * by - uy * bx);

  ux += (vy * bz - vz * by) * bb + ex;
  uy += (vz * bx - vx * bz) * bb + ey;
  uz += (vx * by - vy * bx) * bb + ez;
}

template <int Pusher>
__device__ __forceinline__ void push_momentum(double& ux, double& uy, double& uz, double ex,
                                              double ey, double ez, double bx, double by,
                                              double bz, double cc)
{
  if (Pusher == PICNIX_PUSHER_BORIS) {
    push_boris(ux, uy, uz, ex, ey, ez, bx, by, bz, cc);
  } else if (Pusher == PICNIX_PUSHER_VAY) {
    push_vay(ux, uy, uz, ex, ey, ez, bx, by, bz, cc);
  } else {
    push_higuera_cary(ux, uy, uz, ex, ey, ez, bx, by, bz, cc);
  }
}

// position update, pic/engine/position.hpp:117-130
__device__ __forceinline__ void push_position(double& x, double& y, double& z, double ux,
                                              double uy, double uz, double rc, double delt)
{
  const double gm = sqrt(1 + (ux * ux + uy * uy + uz * uz) * rc * rc);
  const double dt = delt / gm;
  x += ux * dt;
  y += uy * dt;
  z += uz * dt;
}

// Physical boundary of the problem for a particle that has just been moved: the set_boundary_particle
// hooks of the reference's examples, called between the position push and the cell count
// (pic/pic_engine.hpp:292-303).  A conducting wall reflects specularly (example/mrx/main.cpp:352-382),
// the shock tube's wall reverses the whole momentum (example/shock/main.cpp:419-433).
__device__ __forceinline__ void apply_particle_bc(const Geom& g, double& x, double& y, double& z, double& ux,
                                                  double& uy, double& uz)
{
  if (!g.any_particle_bc)
    return;
  double* pos[3] = {&z, &y, &x};
  double* mom[3] = {&uz, &uy, &ux};
#pragma unroll
  for (int axis = 0; axis < 3; axis++) {
#pragma unroll
    for (int side = 0; side < 2; side++) {
      const int kind = g.bc_kind[axis][side];
      if (kind != PICNIX_BC_CONDUCTING && kind != PICNIX_BC_WALL)
        continue;
      const double lim = g.glim[axis][side];
      const bool   out = side == 0 ? *pos[axis] < lim : *pos[axis] >= lim;
      if (out) {
        *pos[axis] = -*pos[axis] + 2 * lim;
        if (kind == PICNIX_BC_CONDUCTING) {
          *mom[axis] = -*mom[axis];
        } else {
          ux = -ux;
          uy = -uy;
          uz = -uz;
        }
      }
    }
  }
}

//
// Cell key of a particle in the geometry of `chunk` (XtensorParticle::count):
// flat index with strides (fsz, fsy, 1) or Ng when the particle is outside the chunk.
// lim = clim + chunk*6 : zmin,zmax,ymin,ymax,xmin,xmax
//
__device__ __forceinline__ int cell_key(const Geom& g, const double* __restrict__ lim, double x,
                                        double y, double z)
{
  const double half = 0.5 * g.is_odd;
  int          ix   = g.has_dim[2] ? digitize(x, lim[4] - half * g.del[2], 1 / g.del[2]) : 0;
  int          iy   = g.has_dim[1] ? digitize(y, lim[2] - half * g.del[1], 1 / g.del[1]) : 0;
  int          iz   = g.has_dim[0] ? digitize(z, lim[0] - half * g.del[0], 1 / g.del[0]) : 0;
  int          ii   = iz * g.fsz + iy * g.fsy + ix;
  ii = (g.has_dim[2] && (x < lim[4] || x >= lim[5])) ? g.Ng : ii;
  ii = (g.has_dim[1] && (y < lim[2] || y >= lim[3])) ? g.Ng : ii;
  ii = (g.has_dim[0] && (z < lim[0] || z >= lim[1])) ? g.Ng : ii;
  return ii;
}

// direction code of XtensorHaloParticle3D::pre_pack (nix/xtensor_halo3d.hpp:233-235):
// 9*iz + 3*iy + ix with i* in {0,1,2}; 13 means "stays"
__device__ __forceinline__ int direction_code(const Geom& g, const double* __restrict__ lim,
                                              double x, double y, double z)
{
  int ix = g.has_dim[2] ? (x >= lim[5]) - (x < lim[4]) + 1 : 1;
  int iy = g.has_dim[1] ? (y >= lim[3]) - (y < lim[2]) + 1 : 1;
  int iz = g.has_dim[0] ? (z >= lim[1]) - (z < lim[0]) + 1 : 1;
  return 9 * iz + 3 * iy + ix;
}

// periodic wrap of a received particle, XtensorParticle::set_boundary_periodic
// (nix/xtensor_particle.hpp:359-376)
__device__ __forceinline__ void wrap_periodic(const Geom& g, double& x, double& y, double& z)
{
  const double X = g.has_dim[2] * (g.glim[2][1] - g.glim[2][0]);
  const double Y = g.has_dim[1] * (g.glim[1][1] - g.glim[1][0]);
  const double Z = g.has_dim[0] * (g.glim[0][1] - g.glim[0][0]);
  x += (x < g.glim[2][0]) * X - (x >= g.glim[2][1]) * X;
  y += (y < g.glim[1][0]) * Y - (y >= g.glim[1][1]) * Y;
  z += (z < g.glim[0][0]) * Z - (z >= g.glim[0][1]) * Z;
}

} // namespace picnix

#endif
