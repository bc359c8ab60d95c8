// -*- C++ -*-
// Row-owner push + Esirkepov deposit for 3-D, 2nd-order shapes (the BASELINE "T3D" configuration).
//
// Why: one fp64 atomic per stencil value (up to 5^3 x 4 per particle) is two orders of magnitude
// too slow -- shared-memory fp64 atomics are CAS loops on sm_100a and every L2 reduction costs an
// LSU slot.  So nothing in the inner loop is atomic.  Instead
//
//   * a block owns WARPS consecutive rows (fixed z, consecutive y) of RX cells; the E/B values all
//     of its particles can touch ((RX+3) x (WARPS+3) x 4 points x 6 components, 14.8 KB) are staged
//     ONCE in shared memory in the global layout, so every interpolation load is an LDS with an
//     immediate offset (no address arithmetic, no L1/L2 latency);
//   * each warp walks the cell-sorted particles of its row segment, 32 at a time:
//       phase 1 (thread per particle): interpolate, push momentum and position, cell key +
//         histogram, then the 1-D Esirkepov factors of the particle on a 4-slot WINDOW per axis --
//         for a 2nd-order shape and |move| < 1 cell the old and new weights together never span
//         more than 4 of the 5 stencil slots (nix/esirkepov.hpp:240-275: the new weights are the
//         old stencil shifted by -1/0/+1) -- written as one 54-double record to shared memory;
//       phase 2 (thread per stencil point): each half-warp takes one record; lane (a,b) owns the
//         window points (a,b,*) and accumulates rho/Jx/Jy/Jz in REGISTERS with one FMA per value
//         (the value is never materialised), exactly the sums of nix/esirkepov.hpp:154-237:
//             rho[z][y][x]   += (S1z[z] S1y[y]) * q S1x[x]
//             Jx [z][y][x+1] += W(y,z) * prefix_x DSx,  W = (S0y+DSy/2) S0z + (S0y/2+DSy/3) DSz
//             (Jy, Jz by cyclic permutation; the factor -q dx/dt is folded into the prefix sums)
//   * when the cell changes the 13 registers of a lane are added -- plain loads/stores, the lanes
//     own distinct points -- into the warp's PRIVATE (RX+4) x 5 x 5 x 4 tile in shared memory;
//   * at the end of the row segment the tile goes to global uj with one fp64 reduction per non-zero
//     tile value (about 2.3 per particle at 64 ppc instead of ~170).
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

constexpr int RX      = 8;           // cells per row segment
constexpr int WARPS   = 4;           // rows (consecutive y) per block
constexpr int THREADS = WARPS * 32;

// field tile: points x in [jx0-1, jx0+RX+1], y in [jy0-1, jy0+WARPS+1], z in [jz-1, jz+2]
constexpr int FX    = RX + 3;
constexpr int FY    = WARPS + 3;
constexpr int FZ    = 4;
constexpr int FROW  = FX * 6;        // doubles per (z, y) row, global layout [x][6]
// z-slab stride.  Lanes of a warp read the same stencil point shifted by hx (12 words), hy (2*FROW
// = 132 = 4 mod 32 words) and hz (one slab); with the natural slab of FY*FROW doubles (28 mod 32
// words) the (hy, hz) = (1, 1) and (0, 0) variants of Bx fall on the same bank.  +6 doubles makes the
// slab 8 mod 32 words: {0,4,8,12} + {0,12} are all distinct bank pairs.
constexpr int FSLAB = FY * FROW + 6;
constexpr int FTILE = FZ * FSLAB;

// current tile of a warp: one array per component, [z][y][x] with x contiguous.  The strides make
// every read-modify-write of flush() conflict free for the 16 lanes (a, b) of a half-warp (two
// words per lane, so a*Sa + b*Sb must be distinct mod 16):
//   rho, Jx: a -> z, b -> y   SZR = 65 (1 mod 16),  SYT = 12:  a + 12 b
//   Jy     : a -> z, b -> x   SZT = 60 (12 mod 16), x = 1   :  12 a + b
//   Jz     : a -> y, b -> x   SYT = 12,             x = 1   :  12 a + b
constexpr int XS   = RX + 4;         // extent in x (stencil reaches -2..+2)
constexpr int SYT  = XS;             // stride of y, all components
constexpr int SZR  = 5 * SYT + 5;    // stride of z for rho and Jx
constexpr int SZT  = 5 * SYT;        // stride of z for Jy and Jz
constexpr int T_RHO = 0, T_JX = 5 * SZR, T_JY = 10 * SZR, T_JZ = 10 * SZR + 5 * SZT;
constexpr int TILE = 10 * SZR + 10 * SZT;
__host__ __device__ constexpr int tile_base(int comp) { return comp == 0 ? T_RHO : (comp == 1 ? T_JX : (comp == 2 ? T_JY : T_JZ)); }
__host__ __device__ constexpr int tile_sz(int comp) { return comp < 2 ? SZR : SZT; }

// staged particle record: 54 doubles = 27 x 16 B (odd multiple: conflict-free 128-bit stores)
constexpr int REC    = 54;
constexpr int O_QS1X = 0;   // q * S1x[4]
constexpr int O_P    = 4;   // Px[3], Py[3], Pz[3]: -q d/dt * prefix sums of DS (window idx 1..3), pad
constexpr int O_A1   = 14;  // [a] -> (S1z[a], S0z[a])
constexpr int O_A2   = 22;  // [a] -> (DSz[a], S0y[a])
constexpr int O_B1   = 30;  // [b] -> (S1y[b], AY[b])     AY = S0y + DSy/2
constexpr int O_B2   = 38;  // [b] -> (BY[b],  AX[b])     BY = S0y/2 + DSy/3, AX = S0x + DSx/2
constexpr int O_C    = 46;  // [i] -> (DSy[i], BX[i])     BX = S0x/2 + DSx/3

struct WarpSmem {
  double stg[32 * REC];
  double tile[TILE];
  int    info[32];
  int    pbuf[32]; // lazy sort: permutation entries of the next batch, filled by cp.async
};

// asynchronous 4-byte global -> shared copy: no register, hence no scoreboard dependency between the
// permutation entry and the particle loads it feeds until the explicit wait
__device__ __forceinline__ void cp_async_i32(int* smem, const int* gmem)
{
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
// 16-byte variant (L2 only): the field tile goes global -> shared without passing through registers
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem)
{
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait()
{
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

constexpr size_t SMEM_BYTES = sizeof(double) * FTILE + sizeof(WarpSmem) * WARPS;

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

  store2(rec + O_QS1X + 0, q * fx.S1[0], q * fx.S1[1]);
  store2(rec + O_QS1X + 2, q * fx.S1[2], q * fx.S1[3]);

  const double px0 = fx.DS[0], px1 = px0 + fx.DS[1], px2 = px1 + fx.DS[2];
  const double py0 = fy.DS[0], py1 = py0 + fy.DS[1], py2 = py1 + fy.DS[2];
  const double pz0 = fz.DS[0], pz1 = pz0 + fz.DS[1], pz2 = pz1 + fz.DS[2];
  store2(rec + O_P + 0, cx * px0, cx * px1);
  store2(rec + O_P + 2, cx * px2, cy * py0);
  store2(rec + O_P + 4, cy * py1, cy * py2);
  store2(rec + O_P + 6, cz * pz0, cz * pz1);
  store2(rec + O_P + 8, cz * pz2, 0.0);

#pragma unroll
  for (int k = 0; k < 4; k++) {
    store2(rec + O_A1 + 2 * k, fz.S1[k], fz.S0[k]);
    store2(rec + O_A2 + 2 * k, fz.DS[k], fy.S0[k]);
    store2(rec + O_B1 + 2 * k, fy.S1[k], fy.S0[k] + A * fy.DS[k]);
    store2(rec + O_B2 + 2 * k, A * fy.S0[k] + B * fy.DS[k], fx.S0[k] + A * fx.DS[k]);
    store2(rec + O_C + 2 * k, fy.DS[k], A * fx.S0[k] + B * fx.DS[k]);
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

// phase 2 body: contributions of one staged particle to the points owned by lane (a, b);
// 13 16-byte shared loads, 7 + 13 fp64 operations
__device__ __forceinline__ void accumulate(Acc& acc, const double* __restrict__ rec, int a, int b)
{
  const double2* r2 = reinterpret_cast<const double2*>(rec);
  const double2  qa = r2[O_QS1X / 2 + 0], qb = r2[O_QS1X / 2 + 1];
  const double2  p0 = r2[O_P / 2 + 0], p1 = r2[O_P / 2 + 1], p2 = r2[O_P / 2 + 2],
                p3 = r2[O_P / 2 + 3], p4 = r2[O_P / 2 + 4];
  const double2 a1 = r2[O_A1 / 2 + a]; // S1z[a], S0z[a]
  const double2 a2 = r2[O_A2 / 2 + a]; // DSz[a], S0y[a]
  const double2 ca = r2[O_C / 2 + a];  // DSy[a], -
  const double2 b1 = r2[O_B1 / 2 + b]; // S1y[b], AY[b]
  const double2 b2 = r2[O_B2 / 2 + b]; // BY[b],  AX[b]
  const double2 cb = r2[O_C / 2 + b];  // -,      BX[b]

  const double c   = a1.x * b1.x;
  const double wyz = b1.y * a1.y + b2.x * a2.x; // AY[b] S0z[a] + BY[b] DSz[a]   (jz=a, jy=b)
  const double wzx = b2.y * a1.y + cb.y * a2.x; // AX[b] S0z[a] + BX[b] DSz[a]   (jz=a, jx=b)
  const double wxy = b2.y * a2.y + cb.y * ca.x; // AX[b] S0y[a] + BX[b] DSy[a]   (jy=a, jx=b)

  acc.rho[0] += c * qa.x;
  acc.rho[1] += c * qa.y;
  acc.rho[2] += c * qb.x;
  acc.rho[3] += c * qb.y;
  acc.jx[0] += wyz * p0.x;
  acc.jx[1] += wyz * p0.y;
  acc.jx[2] += wyz * p1.x;
  acc.jy[0] += wzx * p1.y;
  acc.jy[1] += wzx * p2.x;
  acc.jy[2] += wzx * p2.y;
  acc.jz[0] += wxy * p3.x;
  acc.jz[1] += wxy * p3.y;
  acc.jz[2] += wxy * p4.x;
}

// add a lane's accumulators into the warp tile; (wz,wy,wx) window offset, jx cell in the segment
__device__ __forceinline__ void flush(double* __restrict__ tile, const Acc& acc, int a, int b,
                                      int jx, int wx, int wy, int wz)
{
  // rho and Jx: point (z = wz+a, y = wy+b, x = jx+wx+k)
  double* p = tile + (wz + a) * SZR + (wy + b) * SYT + (jx + wx);
#pragma unroll
  for (int k = 0; k < 4; k++)
    p[T_RHO + k] += acc.rho[k];
#pragma unroll
  for (int k = 0; k < 3; k++)
    p[T_JX + k + 1] += acc.jx[k];
  // Jy: point (z = wz+a, y = wy+k+1, x = jx+wx+b)
  double* py = tile + T_JY + (wz + a) * SZT + wy * SYT + (jx + wx + b);
#pragma unroll
  for (int k = 0; k < 3; k++)
    py[(k + 1) * SYT] += acc.jy[k];
  // Jz: point (z = wz+k+1, y = wy+a, x = jx+wx+b)
  double* pz = tile + T_JZ + wz * SZT + (wy + a) * SYT + (jx + wx + b);
#pragma unroll
  for (int k = 0; k < 3; k++)
    pz[(k + 1) * SZT] += acc.jz[k];
}

// one staged particle straight into the warp tile (no register accumulators): used for the few
// particles whose window differs from the run being accumulated, where accumulate() into a
// temporary followed by flush() would move every value twice
__device__ __forceinline__ void deposit_direct(double* __restrict__ tile, const double* __restrict__ rec,
                                               int a, int b, int jx, int wx, int wy, int wz)
{
  const double2* r2 = reinterpret_cast<const double2*>(rec);
  const double2  qa = r2[O_QS1X / 2 + 0], qb = r2[O_QS1X / 2 + 1];
  const double2  p0 = r2[O_P / 2 + 0], p1 = r2[O_P / 2 + 1], p2 = r2[O_P / 2 + 2],
                p3 = r2[O_P / 2 + 3], p4 = r2[O_P / 2 + 4];
  const double2 a1 = r2[O_A1 / 2 + a];
  const double2 a2 = r2[O_A2 / 2 + a];
  const double2 ca = r2[O_C / 2 + a];
  const double2 b1 = r2[O_B1 / 2 + b];
  const double2 b2 = r2[O_B2 / 2 + b];
  const double2 cb = r2[O_C / 2 + b];

  const double c   = a1.x * b1.x;
  const double wyz = b1.y * a1.y + b2.x * a2.x;
  const double wzx = b2.y * a1.y + cb.y * a2.x;
  const double wxy = b2.y * a2.y + cb.y * ca.x;

  double* p = tile + (wz + a) * SZR + (wy + b) * SYT + (jx + wx);
  p[T_RHO + 0] += c * qa.x;
  p[T_RHO + 1] += c * qa.y;
  p[T_RHO + 2] += c * qb.x;
  p[T_RHO + 3] += c * qb.y;
  p[T_JX + 1] += wyz * p0.x;
  p[T_JX + 2] += wyz * p0.y;
  p[T_JX + 3] += wyz * p1.x;
  double* py = tile + T_JY + (wz + a) * SZT + wy * SYT + (jx + wx + b);
  py[1 * SYT] += wzx * p1.y;
  py[2 * SYT] += wzx * p2.x;
  py[3 * SYT] += wzx * p2.y;
  double* pz = tile + T_JZ + wz * SZT + (wy + a) * SYT + (jx + wx + b);
  pz[1 * SZT] += wxy * p3.x;
  pz[2 * SZT] += wxy * p3.y;
  pz[3 * SZT] += wxy * p4.x;
}

} // namespace rowdep
} // namespace picnix

#endif
