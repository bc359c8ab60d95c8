// -*- C++ -*-
// Row-owner push + Esirkepov deposit for 3-D, 2nd-order shapes (the BASELINE "T3D" configuration),
// second formulation.  What bounds this kernel on B200 is the shared-memory data pipe (one 128-byte
// wavefront per clock and SM): round 1 spent 31 wavefronts per particle (profiles/r01_row_kernel_ncu.txt),
// 10 of them on divergent interpolation loads and 13 on the operands of the deposit.  This version
// cuts both:
//
//   * ONE particle stream per row segment.  The species of a cell are walked back to back
//     (cell-major, species-minor), and every cell starts at an even slot of the stream (at most one
//     idle lane per cell).  Lanes 2k and 2k+1 therefore always hold particles of the SAME cell.
//   * cell-anchored interpolation.  The E/B stencil of a particle is anchored at its CELL, not at its
//     half cell: the three weights of the edge ("half") grid sit in a 4-slot array, shifted by one
//     slot when the particle is right of the cell centre -- exactly what the reference's vector path
//     does (interp::shift_weights, nix/interp.hpp:148-172; pic/engine/velocity.hpp:510-543).  The extra
//     slot carries weight 0, so the sums are bit-identical to the (Order+1)^3 form, but now the load
//     addresses depend on the cell only: both lanes of a pair read the same word and a 64-bit shared
//     load of the warp is ONE wavefront instead of two (tools/micro/lds_patterns.cu, "pairs").
//   * 2-D register tiles in the deposit.  A staged particle is consumed by a half-warp whose lane
//     (c, a) owns component c in {rho, Jx, Jy, Jz} and index a in 0..3 of one axis, i.e. a 4 x 4 patch
//     acc[i][j] of the particle's 4^3 window:
//         rho[z=a][y=i][x=j]   += (S1y[i] S0z[a] + S1y[i] DSz[a]) * q S1x[j]
//         Jx [z=a][y=i][x=j+1] += (AY[i]  S0z[a] + BY[i]  DSz[a]) * Px[j]      AY = S0y + DSy/2
//         Jy [z=a][y=j+1][x=i] += (AX[i]  S0z[a] + BX[i]  DSz[a]) * Py[j]      BY = S0y/2 + DSy/3
//         Jz [z=j+1][y=i][x=a] += (S0y[i] AX[a]  + DSy[i] BX[a])  * Pz[j]      (AX, BX likewise in x)
//     (nix/esirkepov.hpp:154-237 with the factor -q d/dt folded into the prefix sums P).  Every lane runs
//     the SAME code  w[i] = U[i] P + V[i] Q;  acc[i][j] += w[i] R[j]  on tables selected by lane
//     constants: 1 + 4 16-byte loads and 4 8-byte loads feed 24 FMAs (round 1: 13 loads for 20).
//
// The rest is as before: the block stages the E/B tile of its rows once, accumulators are added to
// the warp's private current tile when the cell changes, the tile goes to global uj with one fp64
// reduction per non-zero value at the end of the segment, particles that moved more than one cell go
// to the far-mover list.  Results differ from the reference only by summation order.
#ifndef PICNIX_B200_ROWTILE_CUH
#define PICNIX_B200_ROWTILE_CUH

#include "particle_kernels.cuh"

namespace picnix
{
namespace rowtile
{

constexpr int RX      = 8;           // cells per row segment
constexpr int WARPS   = 4;           // rows (consecutive y) per block
constexpr int THREADS = WARPS * 32;
constexpr int MAXNS   = 4;           // species the merged stream can hold
constexpr int ALIGN   = 2;           // every cell starts at a multiple of ALIGN slots of the stream

// field tile: points x in [jx0-1, jx0+RX+1], y in [jy0-1, jy0+WARPS+1], z in [jz-1, jz+2],
// global layout [z][y][x][6]
constexpr int FX    = RX + 3;
constexpr int FY    = WARPS + 3;
constexpr int FZ    = 4;
constexpr int FROW  = FX * 6;
constexpr int FSLAB = FY * FROW;
constexpr int FTILE = FZ * FSLAB;

// current tile of a warp: [z 5][y 5][x RX+4][component 4], components interleaved like global uj.
// Element index 4 * lin + c with lin = z * SZ + y * SY + x.  The 16 lanes (c, a) of a half-warp add
// their accumulators with 64-bit accesses: the bank pair is (c + 4 * (lin mod 4)) mod 16, and lin mod 4
// runs over all residues with a because a multiplies an ODD stride (SZ for rho/Jx/Jy, 1 for Jz): no
// conflicts whatever the cell or window.
constexpr int XS   = RX + 4;
constexpr int SY   = XS;
constexpr int SZ   = 5 * SY + 1;     // 61: odd
constexpr int TILE = 4 * 5 * SZ;

// staged particle record (doubles): four tables of four 16-byte pairs and R.  The order of the tables and
// the offsets inside R are chosen so that no two of the addresses a warp-level load touches (4 components
// x 2 half-warps; the records of the two half-warps are 46 doubles = 14 bank pairs apart) share a bank:
//    0  ZA[k] = (S0z[k], DSz[k])        8  XA[k] = (AX[k], BX[k])       16  YS[k] = (S0y[k], DSy[k])
//   24  R: q S1x[0..3] | Px[0..2], 0 | Py[0..2] | Pz[0..2]   (offsets 0, 4, 8, 11)
//   38  YA[k] = (AY[k], BY[k])
constexpr int REC  = 46;             // 23 x 16 B: odd multiple -> conflict-free 128-bit stores
constexpr int T_ZA = 0, T_XA = 8, T_YS = 16, T_R = 24, T_YA = 38;

struct WarpSmem {
  double stg[32 * REC];
  double tile[TILE];
  double pfb[7][32];                 // phase space of the next batch, filled by cp.async during phases 1 and 2
  double zero[REC + 2];              // the all-zero record: stands in for a slot that holds no particle of the cell
  double rowc[16];                   // chunk limits and grid points of the row (see rowpush.cu)
  int    info[32];
  int    pbuf[32];                   // lazy sort: permutation entries of the next batch (cp.async)
  // the merged stream of the row segment: entry k = cell * Ns + species covers the stream slots
  // [ent[k].x, ent[k].x + ent[k].z) with the particles ent[k].y, ent[k].y + 1, ... of that species' segment;
  // .w = species | cell << 8; ent[RX * Ns].x = length of the stream (cells padded to ALIGN slots)
  int4   ent[MAXNS * RX + 1];
};

struct BlockSmem {
  double  q[MAXNS];                  // charge
  double  qmdt[MAXNS];               // q/m dt/2
  int64_t off[MAXNS];                // first slot of the (chunk, species) segment
};

static_assert(sizeof(WarpSmem) % 16 == 0, "the records of the next warp must stay 16-byte aligned");

constexpr size_t SMEM_BYTES = sizeof(double) * FTILE + sizeof(BlockSmem) + sizeof(WarpSmem) * WARPS;

__device__ __forceinline__ void cp_async_i32(int* smem, const int* gmem)
{
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_f64(double* smem, const double* gmem)
{
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem)
{
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait()
{
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// info word: bits 0..7 cell index inside the segment, bit 8/9/10 window offset x/y/z (1 = majority
// window that starts at the old cell's slot 1), bit 11 valid
__device__ __forceinline__ int make_info(int jx, int wx, int wy, int wz)
{
  return jx | (wx << 8) | (wy << 9) | (wz << 10) | (1 << 11);
}

// 2nd-order momentum-conserving shape for a normalised offset delta in [-1/2, 1/2]
// (nix/primitives.hpp:266-278)
__device__ __forceinline__ void shape2(double delta, double* s)
{
  const double w1 = 0.5 - delta;
  const double w2 = 0.5 + delta;
  s[0] = 0.50 * w1 * w1;
  s[1] = 0.75 - delta * delta;
  s[2] = 0.50 * w2 * w2;
}

// 1 / x for finite x >= 1 without the special-case branches of the IEEE division: hardware
// approximation + two Newton steps (relative error about one ulp; the parity bar on momenta is 1e-13)
__device__ __forceinline__ double rcp_fast(double x)
{
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}

// Boris push (nix/primitives.hpp:164-189) with the reciprocal square root and the reciprocal taken
// directly instead of through sqrt + two divisions; e* and b* are already multiplied by (q/m) dt / 2
__device__ __forceinline__ void push_boris_fast(double& ux, double& uy, double& uz, double ex, double ey,
                                                double ez, double bx, double by, double bz, double cc)
{
  ux += ex;
  uy += ey;
  uz += ez;
  const double gm = rsqrt(cc * cc + ux * ux + uy * uy + uz * uz);
  bx *= gm;
  by *= gm;
  bz *= gm;
  const double bb = 2.0 * rcp_fast(1.0 + bx * bx + by * by + bz * bz);
  const double vx = ux + (uy * bz - uz * by);
  const double vy = uy + (uz * bx - ux * bz);
  const double vz = uz + (ux * by - uy * bx);
  ux += (vy * bz - vz * by) * bb + ex;
  uy += (vz * bx - vx * bz) * bb + ey;
  uz += (vx * by - vy * bx) * bb + ez;
}

// position update (pic/engine/position.hpp:117-130): x += u dt / gamma
__device__ __forceinline__ void push_position_fast(double& x, double& y, double& z, double ux, double uy,
                                                   double uz, double rc, double delt)
{
  const double dt = delt * rsqrt(1 + (ux * ux + uy * uy + uz * uz) * rc * rc);
  x += ux * dt;
  y += uy * dt;
  z += uz * dt;
}

// the three weights of the edge grid in the cell-anchored 4-slot array: slot k <-> edge cell-1+k;
// up = the nearest edge is the right one (interp::shift_weights, nix/interp.hpp:148-172)
__device__ __forceinline__ void shift4(const double* h, bool up, double* w4)
{
  w4[0] = up ? 0.0 : h[0];
  w4[1] = up ? h[0] : h[1];
  w4[2] = up ? h[1] : h[2];
  w4[3] = up ? h[2] : 0.0;
}

// tensor-product interpolation on the shared field tile, x innermost (nix/interp.hpp:94-113); p points
// at the cell-anchored first stencil point of the wanted component
template <int NZ, int NY, int NX>
__device__ __forceinline__ double interp_cell(const double* __restrict__ p, const double* wz,
                                              const double* wy, const double* wx)
{
  double rz = 0;
#pragma unroll
  for (int jz = 0; jz < NZ; jz++) {
    double ry = 0;
#pragma unroll
    for (int jy = 0; jy < NY; jy++) {
      double rx = 0;
#pragma unroll
      for (int jx = 0; jx < NX; jx++)
        rx += p[jz * FSLAB + jy * FROW + jx * 6] * wx[jx];
      ry += rx * wy[jy];
    }
    rz += ry * wz[jz];
  }
  return rz;
}

// old/new weights of one axis on the 4-slot window; s0/s1 are the 3 weights around the old/new cell
struct AxisFactors {
  double S0[4], S1[4], DS[4];
  int    w; // window offset: 1 = slots 1..4 of the 5-slot stencil, 0 = slots 0..3
};

__device__ __forceinline__ AxisFactors window_factors(const double* s0, const double* s1, int sh)
{
  AxisFactors f;
  f.w           = sh < 0 ? 0 : 1;
  const bool w1 = sh >= 0;
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

__device__ __forceinline__ void store2(double* p, double a, double b)
{
  *reinterpret_cast<double2*>(p) = make_double2(a, b);
}

// phase 1: stage the factors of one particle (lane-private record, 16-byte stores)
__device__ __forceinline__ void stage_particle(double* __restrict__ rec, const AxisFactors& fx,
                                               const AxisFactors& fy, const AxisFactors& fz,
                                               double q, double dxdt, double dydt, double dzdt)
{
  const double A = 1.0 / 2, B = 1.0 / 3;
  const double cx = -q * dxdt, cy = -q * dydt, cz = -q * dzdt;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    store2(rec + T_ZA + 2 * k, fz.S0[k], fz.DS[k]);
    store2(rec + T_XA + 2 * k, fx.S0[k] + A * fx.DS[k], A * fx.S0[k] + B * fx.DS[k]);
    store2(rec + T_YS + 2 * k, fy.S0[k], fy.DS[k]);
    store2(rec + T_YA + 2 * k, fy.S0[k] + A * fy.DS[k], A * fy.S0[k] + B * fy.DS[k]);
  }
  const double px0 = fx.DS[0], px1 = px0 + fx.DS[1], px2 = px1 + fx.DS[2];
  const double py0 = fy.DS[0], py1 = py0 + fy.DS[1], py2 = py1 + fy.DS[2];
  const double pz0 = fz.DS[0], pz1 = pz0 + fz.DS[1], pz2 = pz1 + fz.DS[2];
  store2(rec + T_R + 0, q * fx.S1[0], q * fx.S1[1]);
  store2(rec + T_R + 2, q * fx.S1[2], q * fx.S1[3]);
  store2(rec + T_R + 4, cx * px0, cx * px1);
  store2(rec + T_R + 6, cx * px2, 0.0);
  store2(rec + T_R + 8, cy * py0, cy * py1);
  store2(rec + T_R + 10, cy * py2, cz * pz0);
  store2(rec + T_R + 12, cz * pz1, cz * pz2);
}

// lane constants of phase 2: lane (c, a) = (component, index) inside its half-warp
struct LaneMap {
  int pq;   // record offset of (P, Q)
  int uv;   // record offset of the (U[i], V[i]) table
  int r;    // record offset of R[0..3] (the currents have three prefix values; their R[3] is whatever
            // follows and feeds a column that is never flushed)
  double e; // 1 for rho, 0 for the currents: (P, Q) <- (P + e Q, Q + e P) turns (S0z, DSz) into (S1z, S1z)
  int lin;  // lane part of the tile index (the run adds wz*SZ + wy*SY + jx + wx)
  int si;   // tile stride of i, in elements (already x 4 components)
  int sj;   // tile stride of j
  int c;    // component
};

__device__ __forceinline__ LaneMap lane_map(int lane)
{
  const int a = lane & 3;
  const int c = (lane >> 2) & 3;
  LaneMap   m;
  m.c   = c;
  m.pq  = (c == 3 ? T_XA : T_ZA) + 2 * a;
  m.uv  = c == 1 ? T_YA : (c == 2 ? T_XA : T_YS);
  m.e   = c == 0 ? 1.0 : 0.0;
  m.r   = T_R + (c == 0 ? 0 : (c == 1 ? 4 : (c == 2 ? 8 : 11)));
  m.lin = c == 0 ? a * SZ : (c == 1 ? a * SZ + 1 : (c == 2 ? a * SZ + SY : SZ + a));
  m.si  = 4 * (c == 2 ? 1 : SY);
  m.sj  = 4 * (c == 2 ? SY : (c == 3 ? SZ : 1));
  return m;
}

// the 4 x 4 register patch of a lane
struct Acc {
  double v[4][4];
  __device__ __forceinline__ void clear()
  {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++)
        v[i][j] = 0;
  }
};

struct Operands {
  double w[4], r[4];
};

// the factors of one staged particle as seen by lane (c, a): 5 16-byte + 4 8-byte shared loads.
//   rho: w[i] = S0y[i] S1z[a] + DSy[i] S1z[a] = S1y[i] S1z[a]      (P, Q) = (S0z, DSz)[a] -> (S1z, S1z)[a]
__device__ __forceinline__ Operands load_operands(const double* __restrict__ rec, const LaneMap& m)
{
  Operands      o;
  const double2 pq = *reinterpret_cast<const double2*>(rec + m.pq);
  const double  P = pq.x + m.e * pq.y, Q = pq.y + m.e * pq.x;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const double2 uv = *reinterpret_cast<const double2*>(rec + m.uv + 2 * i);
    o.w[i]           = uv.x * P + uv.y * Q;
  }
#pragma unroll
  for (int j = 0; j < 4; j++)
    o.r[j] = rec[m.r + j];
  return o;
}

// phase 2 body: contributions of one staged particle to the patch of lane (c, a)
__device__ __forceinline__ void accumulate(Acc& acc, const double* __restrict__ rec, const LaneMap& m)
{
  const Operands o = load_operands(rec, m);
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
      acc.v[i][j] += o.w[i] * o.r[j];
}

// Both half-warps hold a patch of the SAME cell and window (run = wz*SZ + wy*SY + jx + wx).  They first
// add the two patches lane-wise -- each half keeps two of the four rows i and receives the partner's
// contribution to them by shuffle -- and then every lane adds its two rows into the warp tile: all 32
// lanes busy, no serialisation of the halves.  The current components only have three prefix values (j < 3).
__device__ __forceinline__ void flush(double* __restrict__ tile, const Acc& acc, const LaneMap& m, int run,
                                      int half)
{
  double* p = tile + 4 * (m.lin + run) + m.c + 2 * half * m.si;
#pragma unroll
  for (int ii = 0; ii < 2; ii++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const double give = half ? acc.v[ii][j] : acc.v[2 + ii][j];     // the partner's row
      const double keep = half ? acc.v[2 + ii][j] : acc.v[ii][j];     // row i = 2 * half + ii
      const double sum  = keep + __shfl_xor_sync(0xffffffffu, give, 16);
      if (j < 3 || m.c == 0)
        p[ii * m.si + j * m.sj] += sum;
    }
  }
}

// one staged particle straight into the warp tile (no register accumulators), the whole warp on it:
// half-warp h adds rows i = 2h, 2h+1 of every lane's patch.  For the few particles whose window
// differs from the run being accumulated.
__device__ __forceinline__ void deposit_direct(double* __restrict__ tile, const double* __restrict__ rec,
                                               const LaneMap& m, int run, int half)
{
  const double2 pq = *reinterpret_cast<const double2*>(rec + m.pq);
  const double  P = pq.x + m.e * pq.y, Q = pq.y + m.e * pq.x;
  double        r[4];
#pragma unroll
  for (int j = 0; j < 4; j++)
    r[j] = rec[m.r + j];
  double* p = tile + 4 * (m.lin + run) + m.c + 2 * half * m.si;
#pragma unroll
  for (int ii = 0; ii < 2; ii++) {
    const double2 uv = *reinterpret_cast<const double2*>(rec + m.uv + 2 * (2 * half + ii));
    const double  w  = uv.x * P + uv.y * Q;
#pragma unroll
    for (int j = 0; j < 3; j++)
      p[ii * m.si + j * m.sj] += w * r[j];
    if (m.c == 0)
      p[ii * m.si + 3 * m.sj] += w * r[3];
  }
}

__device__ __forceinline__ int run_index(int info)
{
  const int jx = info & 0xff;
  const int wx = (info >> 8) & 1, wy = (info >> 9) & 1, wz = (info >> 10) & 1;
  return wz * SZ + wy * SY + jx + wx;
}

} // namespace rowtile
} // namespace picnix

#endif
