// -*- C++ -*-
// Esirkepov deposit as per-cell sums of outer products on the FP64 MMA unit (DMMA.8x8x4).
//
// For the particles p of one cell, on the 4-slot windows of rowdeposit.cuh, the deposit is
//     rho[z][y][x]   = sum_p  c_p[z,y]   * qS1x_p[x]        c   = S1z (x) S1y
//     Jx [z][y][x+1] = sum_p  wyz_p[z,y] * Px_p[x]          wyz = AY S0z + BY DSz
//     Jy [z][y+1][x] = sum_p  wzx_p[z,x] * Py_p[y]          wzx = AX S0z + BX DSz
//     Jz [z+1][y][x] = sum_p  wxy_p[y,x] * Pz_p[z]          wxy = AX S0y + BX DSy
// i.e. four products  A (16 x P) * B (P x 4|3)  with the particle index as the contraction index
// (nix/esirkepov.hpp:154-237 written as sums over particles).  The scalar kernel delivers every
// operand of every FMA through shared memory (26 doubles per lane for 13 FMAs) and is bound by the
// shared-memory data pipe; mma.sync.m8n8k4.f64 takes ONE A and ONE B element per lane for 256 FMAs.
// On B200 DMMA and DFMA share the same 64 FMA/clk/SM pipe (tools/micro/fp64_pipes.cu), so this buys
// no FP64 rate -- it removes operand traffic and issue slots:
//
//   * phase 1 (thread per particle, as before) expands its 1-D factors into the A vectors and
//     stages them element-major in shared memory, in two rounds so that the buffer stays 11.5 KB:
//     round 0 = (c, wyz | qS1x, Px), round 1 = (wzx, wxy | Py, Pz);
//   * phase 2: for each group of 4 staged particles (the k index of the MMA) a lane loads its
//     A elements (rows g and g+8 of both vectors) and its B element: 5 conflict-free 64-bit loads
//     feed 4 DMMAs = 1024 FMAs; B columns are [qS1x | Px | 0] resp. [Py 0 | Pz 0], so the useful
//     results sit in columns 0-3 of the first and 4-6 of the second accumulator pair;
//   * particles of another cell or window inside a group are masked out through a zero B element
//     and taken in a second pass; when the cell changes the accumulators (fragment layout:
//     row = lane/4 (+8), columns 2*(lane%4)+{0,1}) go straight to global uj with one fp64 reduction
//     per value -- there is no per-warp current tile any more, which frees 10 KB of shared memory
//     per warp and lets three blocks instead of two share an SM.
//
// MEASURED (T3D, B200, profiles/r01_row_mma_ncu.txt): correct (same parity tests), but 36.3 ms against
// 20.9 ms of the scalar row kernel.  The run bookkeeping (a group of four particles spans a cell
// change or a shifted window 30 % of the time), the 8 % of particles that need their own flush and
// the eight reductions per lane and flush add more instructions (13.6 G vs 9.5 G warp instructions)
// than the operand traffic they save (3.5 G vs 4.2 G shared wavefronts), and the padded columns
// double the FP64 work of the deposit.  It stays selectable (option "deposit_mma") as the evidence
// behind "no tensor cores on this path"; the scalar kernel is the product path.
#ifndef PICNIX_B200_ROWMMA_CUH
#define PICNIX_B200_ROWMMA_CUH

#include "rowdeposit.cuh"

namespace picnix
{
namespace rowmma
{

using rowdep::AxisFactors;

constexpr int NELEM = 40;  // staged elements per particle and round: A0[16], A1[16], B[8]
constexpr int RS    = 36;  // row stride in doubles (32 particles + 4): 8 mod 32 words, so the 16 lanes
                           // (g = 0..3, k = 0..3) of a half-warp hit 16 distinct bank pairs
constexpr int E_A0 = 0, E_A1 = 16, E_B = 32;

struct WarpSmem {
  double stg[NELEM * RS];
  int    info[32];
};

constexpr size_t SMEM_BYTES = sizeof(double) * rowdep::FTILE + sizeof(WarpSmem) * rowdep::WARPS;

// what phase 1 keeps of a particle between the two staging rounds
struct Factors {
  double S1z[4], S0z[4], DSz[4];
  double S1y[4], S0y[4], DSy[4], AY[4], BY[4];
  double AX[4], BX[4];
  double qS1x[4], Px[3], Py[3], Pz[3];
};

__device__ __forceinline__ Factors make_factors(const AxisFactors& fx, const AxisFactors& fy,
                                                const AxisFactors& fz, double q, double dxdt,
                                                double dydt, double dzdt)
{
  Factors      f;
  const double A = 1.0 / 2, B = 1.0 / 3;
  const double cx = -q * dxdt, cy = -q * dydt, cz = -q * dzdt;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    f.S1z[k]  = fz.S1[k];
    f.S0z[k]  = fz.S0[k];
    f.DSz[k]  = fz.DS[k];
    f.S1y[k]  = fy.S1[k];
    f.S0y[k]  = fy.S0[k];
    f.DSy[k]  = fy.DS[k];
    f.AY[k]   = fy.S0[k] + A * fy.DS[k];
    f.BY[k]   = A * fy.S0[k] + B * fy.DS[k];
    f.AX[k]   = fx.S0[k] + A * fx.DS[k];
    f.BX[k]   = A * fx.S0[k] + B * fx.DS[k];
    f.qS1x[k] = q * fx.S1[k];
  }
  const double px0 = fx.DS[0], px1 = px0 + fx.DS[1], px2 = px1 + fx.DS[2];
  const double py0 = fy.DS[0], py1 = py0 + fy.DS[1], py2 = py1 + fy.DS[2];
  const double pz0 = fz.DS[0], pz1 = pz0 + fz.DS[1], pz2 = pz1 + fz.DS[2];
  f.Px[0] = cx * px0, f.Px[1] = cx * px1, f.Px[2] = cx * px2;
  f.Py[0] = cy * py0, f.Py[1] = cy * py1, f.Py[2] = cy * py2;
  f.Pz[0] = cz * pz0, f.Pz[1] = cz * pz1, f.Pz[2] = cz * pz2;
  return f;
}

// stage round 0 (rho, Jx) or round 1 (Jy, Jz) of particle `p` (= lane): element-major, so the 32
// lanes of a store write 32 consecutive doubles
template <int ROUND>
__device__ __forceinline__ void stage_round(double* __restrict__ stg, int p, const Factors& f)
{
  double* s = stg + p;
  if (ROUND == 0) {
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) {
        s[(E_A0 + a * 4 + b) * RS] = f.S1z[a] * f.S1y[b];
        s[(E_A1 + a * 4 + b) * RS] = f.AY[b] * f.S0z[a] + f.BY[b] * f.DSz[a];
      }
#pragma unroll
    for (int k = 0; k < 4; k++)
      s[(E_B + k) * RS] = f.qS1x[k];
#pragma unroll
    for (int k = 0; k < 3; k++)
      s[(E_B + 4 + k) * RS] = f.Px[k];
    s[(E_B + 7) * RS] = 0.0;
  } else {
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) {
        s[(E_A0 + a * 4 + b) * RS] = f.AX[b] * f.S0z[a] + f.BX[b] * f.DSz[a];
        s[(E_A1 + a * 4 + b) * RS] = f.AX[b] * f.S0y[a] + f.BX[b] * f.DSy[a];
      }
#pragma unroll
    for (int k = 0; k < 3; k++) {
      s[(E_B + k) * RS]     = f.Py[k];
      s[(E_B + 4 + k) * RS] = f.Pz[k];
    }
    s[(E_B + 3) * RS] = 0.0;
    s[(E_B + 7) * RS] = 0.0;
  }
}

// D(8x8) += A(8x4) * B(4x8), fp64.  Lane l: A[l/4][l%4], B[l%4][l/4], D[l/4][2*(l%4) + {0,1}]
__device__ __forceinline__ void dmma(double (&d)[2], double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d[0]), "+d"(d[1])
               : "d"(a), "d"(b));
}

// accumulators of one round: [vector 0|1][row tile 0|1][column 0|1 of the lane's pair]
struct Frag {
  double v[2][2][2];
  __device__ __forceinline__ void clear()
  {
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int t = 0; t < 2; t++)
        v[i][t][0] = v[i][t][1] = 0.0;
  }
};

// one group of 4 staged particles (kg) into the accumulators; `sel` is false on the lanes whose
// particle (k) does not belong to the run being accumulated
__device__ __forceinline__ void mma_group(Frag& acc, const double* __restrict__ stg, int kg, int g,
                                          int k, bool sel)
{
  const double* s  = stg + kg * 4 + k;
  const double  a00 = s[(E_A0 + g) * RS], a01 = s[(E_A0 + 8 + g) * RS];
  const double  a10 = s[(E_A1 + g) * RS], a11 = s[(E_A1 + 8 + g) * RS];
  const double  b   = s[(E_B + g) * RS];
  // a deselected slot may hold stale bytes of an earlier batch: zero BOTH operands (0 * NaN = NaN)
  const double  z   = 0.0;
  dmma(acc.v[0][0], sel ? a00 : z, sel ? b : z);
  dmma(acc.v[0][1], sel ? a01 : z, sel ? b : z);
  dmma(acc.v[1][0], sel ? a10 : z, sel ? b : z);
  dmma(acc.v[1][1], sel ? a11 : z, sel ? b : z);
}

// accumulators -> global uj.  (oz, oy, ox): global index of window slot 0 of the cell the run
// belongs to.  Round 0: vector 0 = rho (columns 0-3 = x slot), vector 1 = Jx (columns 4-6 = x slot - 3).
// Round 1: vector 0 = Jy (columns 0-2 = y slot - 1; rows (z, x)), vector 1 = Jz (columns 4-6 =
// z slot - 3; rows (y, x)).
template <int ROUND>
__device__ __forceinline__ void flush_global(double* __restrict__ uj, int My, int Mx, const Frag& acc,
                                             int g, int k, int oz, int oy, int ox)
{
#pragma unroll
  for (int t = 0; t < 2; t++) {
    const int a = 2 * t + (g >> 2), b = g & 3;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int col = 2 * k + j;
      int       z, y, x, comp;
      double    v;
      if (ROUND == 0) {
        if (col < 4) { // rho[z=a][y=b][x=col]
          v = acc.v[0][t][j], z = a, y = b, x = col, comp = 0;
        } else {       // Jx[z=a][y=b][x=col-3]
          v = acc.v[1][t][j], z = a, y = b, x = col - 3, comp = 1;
        }
      } else {
        if (col < 4) { // Jy[z=a][y=col+1][x=b]
          v = acc.v[0][t][j], z = a, y = col + 1, x = b, comp = 2;
        } else {       // Jz[z=col-3][y=a][x=b]
          v = acc.v[1][t][j], z = col - 3, y = a, x = b, comp = 3;
        }
      }
      if (v != 0.0 && col != 7 && !(ROUND == 1 && col == 3))
        atomicAdd(uj + ((int64_t)((oz + z) * My + (oy + y)) * Mx + (ox + x)) * 4 + comp, v);
    }
  }
}

} // namespace rowmma
} // namespace picnix

#endif
