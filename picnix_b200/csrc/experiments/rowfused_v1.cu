// -*- C++ -*-
// NOT THE PRODUCT PATH.  Round-1 row-owner kernel and the FP64-MMA deposit experiment, kept buildable as
// measured evidence (profiles/r01_row_kernel_ncu.txt, profiles/r01_row_mma_ncu.txt: the MMA deposit is
// 1.74x slower) and as the tiled path for more than rowtile::MAXNS species.  Selected only
// by set_option("row_kernel", 1) / set_option("deposit_mma", 1); the hot kernels are rowpush.cu (3-D)
// and rowpush2d.cu (2-D), dispatched from rowdispatch.cu.
//
// Row-owner kernel (3-D, order 2): see rowdeposit.cuh for the scheme.
//   FUSED = true : phase 1 interpolates the fields, pushes momentum and position, writes the new
//                  state in place and produces the cell key + histogram (K1 + K2 in one pass over
//                  the particles: PicChunk::push_velocity + push_position + deposit_current,
//                  pic/pic_chunk.cpp:491-523)
//   FUSED = false: deposit only, old position from xv, new position from xu (PicChunk::deposit_current)
// One block per (chunk, z, group of WARPS rows in y, x-segment of RX cells), one warp per row; the
// particles of a row segment are contiguous in the cell-sorted arrays: [pindex[key0], pindex[key0+RX]).
#include "rowmma.cuh"

namespace picnix
{

namespace
{

using namespace rowdep;

// chunk-independent constants of the run, computed once on the host
struct RowConst {
  double rd[3];    // 1/dz, 1/dy, 1/dx
  double del[3];   // dz, dy, dx
  double ddt[3];   // dz/dt, dy/dt, dx/dt
  double cc, rc, delt, cfl[3];
};

// Particles that moved more than one cell (never at a Courant-limited time step; the parity tests
// provoke it with large steps) do not fit the 4-slot window.  They are appended to a list and
// deposited by far_kernel with the generic stencil, which keeps that code out of the hot kernel.
__device__ __forceinline__ void defer_far_mover(const DevPtrs& d, int chunk, double q, double x0,
                                                double y0, double z0, double x1, double y1,
                                                double z1)
{
  const int slot = atomicAdd(d.far_count, 1);
  if (slot >= d.far_cap) {
    atomicExch(d.errflag + 3, 1);
    return;
  }
  double* r = d.far_rec + (int64_t)slot * 8;
  r[0] = x0;
  r[1] = y0;
  r[2] = z0;
  r[3] = x1;
  r[4] = y1;
  r[5] = z1;
  r[6] = q;
  r[7] = (double)chunk;
}

__global__ void __launch_bounds__(128) far_kernel(Geom g, DevPtrs d, double delt)
{
  const int n = min(*d.far_count, d.far_cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double* r     = d.far_rec + (int64_t)i * 8;
    const int     chunk = (int)r[7];
    const double* lim   = d.clim + chunk * 6;
    double*       uj    = d.uj + (int64_t)chunk * g.Ng * 4;
    int           bz = 0, by = 0, bx = 0;
    const int     My = g.M[1], Mx = g.M[2];
    auto          add = [&](int kz, int ky, int kx, int k, double v) {
      if (v != 0.0)
        atomicAdd(uj + ((int64_t)((bz + kz) * My + (by + ky)) * Mx + (bx + kx)) * 4 + k, v);
    };
    esirkepov_deposit<3, 2>(g, lim, r[6], delt, r[0], r[1], r[2], r[3], r[4], r[5], bz, by, bx, add);
  }
}

// tensor-product interpolation on the shared field tile, x innermost (nix/interp.hpp:94-113);
// p points at the first stencil point of the wanted component
__device__ __forceinline__ double interp27(const double* __restrict__ p, const double* wz,
                                           const double* wy, const double* wx)
{
  double rz = 0;
#pragma unroll
  for (int jz = 0; jz < 3; jz++) {
    double ry = 0;
#pragma unroll
    for (int jy = 0; jy < 3; jy++) {
      double rx = 0;
#pragma unroll
      for (int jx = 0; jx < 3; jx++)
        rx += p[jz * FSLAB + jy * FROW + jx * 6] * wx[jx];
      ry += rx * wy[jy];
    }
    rz += ry * wz[jz];
  }
  return rz;
}

// PERM (fused only): a lazy sort is pending -- sorted slot j of a segment still sits in slot perm[j] of
// xu; the kernel reads through the permutation and writes the pushed particle (all seven components)
// to slot j of xv, so the reordering costs no pass of its own (the host swaps xu/xv afterwards).
template <bool FUSED, int Pusher, int Interp, bool PERM>
__global__ void __launch_bounds__(THREADS, 2)
row_kernel(Geom g, DevPtrs d, RowConst rc, int c0, int cn, double delt)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double*   ftile = reinterpret_cast<double*>(smem_raw);
  WarpSmem* wsm   = reinterpret_cast<WarpSmem*>(smem_raw + sizeof(double) * FTILE);

  const int      lane = threadIdx.x & 31;
  const int      warp = threadIdx.x >> 5;
  const int      half = lane >> 4;
  // (a, b) of a lane inside its half-warp.  A shared-memory wavefront serves each aligned group of
  // 4 lanes with at most two distinct addresses in the patterns XXYY / XYXY (measured,
  // tools/micro/lds_patterns.cu), so both the a-indexed and the b-indexed record loads of phase 2
  // must take only two values per quad: bit 1 of the lane is the low bit of a, bit 0 the low bit of b.
  const int      a    = ((lane >> 2) & 2) | ((lane >> 1) & 1);
  const int      b    = ((lane >> 1) & 2) | (lane & 1);
  const unsigned FULL = 0xffffffffu;
  WarpSmem*      ws   = wsm + warp;

  // block -> (chunk, z, y group, x segment)
  const int nsegx = g.dims[2] / RX;
  const int nygrp = g.dims[1] / WARPS;
  int       r     = blockIdx.x;
  const int lc    = r / (g.dims[0] * nygrp * nsegx);
  r -= lc * g.dims[0] * nygrp * nsegx;
  const int jz = r / (nygrp * nsegx);
  r -= jz * nygrp * nsegx;
  const int jy0   = (r / nsegx) * WARPS;
  const int jx0   = (r - (r / nsegx) * nsegx) * RX;
  const int jy    = jy0 + warp;
  const int chunk = c0 + lc;

  const double* lim = d.clim + chunk * 6;
  double*       uj  = d.uj + (int64_t)chunk * g.Ng * 4;
  const int     My = g.M[1], Mx = g.M[2];

  // ---- particle range of the first species and its first batch: requested BEFORE the field tile is
  // staged, so that the dependent chain pindex -> permutation -> particle data overlaps the tile loads
  const int key0 = jz * g.fsz + jy * g.fsy + jx0;
  // (inside the species loop the range of the NEXT species is requested at the top of the current
  // one, and its first batch during the last batch of the current one)
  int64_t off;
  int     pb, pe;
  {
    const int  seg0 = chunk * g.Ns;
    const int* pix0 = d.pindex + (int64_t)seg0 * (g.Ng + 1);
    off             = d.seg_off[seg0];
    pb              = pix0[key0];
    pe              = pix0[key0 + RX];
  }
  bool   primed = false; // the first batch of the current species is already in the pf* registers
  double pfx = 0, pfy = 0, pfz = 0, pfux = 0, pfuy = 0, pfuz = 0, pfid = 0;

  if (FUSED && pb + lane < pe) {
    const int64_t i = PERM ? off + d.perm[off + pb + lane] : off + pb + lane;
    if (PERM)
      pfid = d.xu[6 * d.pcap + i];
    pfx    = d.xu[0 * d.pcap + i];
    pfy    = d.xu[1 * d.pcap + i];
    pfz    = d.xu[2 * d.pcap + i];
    pfux   = d.xu[3 * d.pcap + i];
    pfuy   = d.xu[4 * d.pcap + i];
    pfuz   = d.xu[5 * d.pcap + i];
  }
  primed = FUSED && pe > pb;

  // ---- stage the field tile (global layout [z][y][x][6], 16-byte copies) ----
  if (FUSED) {
    const double* uf = d.uf + (int64_t)chunk * g.Ng * 6;
    const int     gz = jz + g.Lb[0] - 1, gy = jy0 + g.Lb[1] - 1, gx = jx0 + g.Lb[2] - 1;
    for (int e = threadIdx.x; e < FZ * FY * (FROW / 2); e += THREADS) {
      const int row = e / (FROW / 2);
      const int col = e - row * (FROW / 2);
      const int tz  = row / FY;
      const int ty  = row - tz * FY;
      // asynchronous copy: the tile arrives while the current tile is being cleared below
      cp_async_16(reinterpret_cast<double2*>(ftile + tz * FSLAB + ty * FROW) + col,
                  reinterpret_cast<const double2*>(uf + ((int64_t)((gz + tz) * My + (gy + ty)) * Mx + gx) * 6) + col);
    }
  }
  for (int i = lane; i < TILE; i += 32)
    ws->tile[i] = 0.0;
  if (FUSED) {
    cp_async_commit_wait();
    __syncthreads();
  } else {
    __syncwarp();
  }

  const double xmin = lim[4], ymin = lim[2], zmin = lim[0];
  const double xmax = lim[5], ymax = lim[3], zmax = lim[1]; // in registers: the key tests run per particle
  const double rdx = rc.rd[2], rdy = rc.rd[1], rdz = rc.rd[0];
  const double dx = rc.del[2], dy = rc.del[1], dz = rc.del[0];
  // cell-centre ("integer") and cell-edge ("half") grid points of this row, pic/engine/velocity.hpp:304-315
  const double yig = ymin + 0.5 * dy + (double)jy * dy;
  const double zig = zmin + 0.5 * dz + (double)jz * dz;
  const double yh0 = ymin + (double)jy * dy, yh1 = ymin + (double)(jy + 1) * dy;
  const double zh0 = zmin + (double)jz * dz, zh1 = zmin + (double)(jz + 1) * dz;
  const double xigrid = xmin + 0.5 * dx, yigrid = ymin + 0.5 * dy, zigrid = zmin + 0.5 * dz;

  for (int is = 0; is < g.Ns; is++) {
    const int seg = chunk * g.Ns + is;
    int64_t   noff = 0;
    int       npb = 0, npe = 0;
    if (is + 1 < g.Ns) {
      const int* pix1 = d.pindex + (int64_t)(seg + 1) * (g.Ng + 1);
      noff            = d.seg_off[seg + 1];
      npb             = pix1[key0];
      npe             = pix1[key0 + RX];
    }
    const double  q    = d.qm[2 * is];
    const double  qmdt = 0.5 * q / d.qm[2 * is + 1] * delt;

    Acc acc;
    acc.clear();
    int curinfo = -1; // info word of the cell the accumulators belong to (-1: none)

    // the six phase-space components of the NEXT batch are requested before phase 2 of the current
    // one, so their HBM latency is hidden behind the accumulation loop
    if (FUSED && !primed && pb + lane < pe) {
      const int64_t i = PERM ? off + d.perm[off + pb + lane] : off + pb + lane;
      if (PERM)
        pfid = d.xu[6 * d.pcap + i];
      pfx  = d.xu[0 * d.pcap + i];
      pfy  = d.xu[1 * d.pcap + i];
      pfz  = d.xu[2 * d.pcap + i];
      pfux = d.xu[3 * d.pcap + i];
      pfuy = d.xu[4 * d.pcap + i];
      pfuz = d.xu[5 * d.pcap + i];
    }

    for (int base = pb; base < pe; base += 32) {
      const int n = min(32, pe - base);
      // the batch whose data is requested after phase 1: the next one of this species, or the first
      // one of the next species
      const bool    last  = base + 32 >= pe;
      const int64_t xoff  = last ? noff : off;
      const int     xbase = last ? npb : base + 32;
      const int     xend  = last ? npe : pe;
      // PERM: its permutation entries travel global -> shared asynchronously while phase 1 runs; they
      // are read (and the particle loads issued) after phase 1
      if (PERM && xbase + lane < xend)
        cp_async_i32(ws->pbuf + lane, d.perm + xoff + xbase + lane);

      // ---------------- phase 1: one particle per lane ----------------
      int inf = 0;
      if (lane < n) {
        const int64_t i = off + base + lane;
        double        x0, y0, z0, x1, y1, z1;
        double        s0x[3], s0y[3], s0z[3];
        int           cx; // old cell in x, relative to the chunk
        if (FUSED) {
          x0        = pfx;
          y0        = pfy;
          z0        = pfz;
          double ux = pfux;
          double uy = pfuy;
          double uz = pfuz;

          // weights on the centre grid (MC or WT) and on the edge grid (MC); the particle is in
          // row (jz, jy) by construction of the sort, only its x cell has to be found
          double wix[3], whx[3], wiy[3], why[3], wiz[3], whz[3];
          cx = digitize(x0, xmin, rdx);
          const double cxf = (double)cx;
          const double dix = (x0 - (xigrid + cxf * dx)) * rdx;
          const double diy = (y0 - yig) * rdy;
          const double diz = (z0 - zig) * rdz;
          shape2(dix, s0x);
          shape2(diy, s0y);
          shape2(diz, s0z);
          if (Interp == PICNIX_INTERP_MC) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
              wix[k] = s0x[k];
              wiy[k] = s0y[k];
              wiz[k] = s0z[k];
            }
          } else {
            shape_wt<2>(x0, xigrid + cxf * dx, rdx, rc.cfl[2], 1 / rc.cfl[2], wix);
            shape_wt<2>(y0, yig, rdy, rc.cfl[1], 1 / rc.cfl[1], wiy);
            shape_wt<2>(z0, zig, rdz, rc.cfl[0], 1 / rc.cfl[0], wiz);
          }
          // nearest cell edge: the one to the right when the particle sits right of the centre
          const int hx = dix >= 0.0, hy = diy >= 0.0, hz = diz >= 0.0;
          shape2((x0 - (xmin + (cxf + (double)hx) * dx)) * rdx, whx);
          shape2((y0 - (hy ? yh1 : yh0)) * rdy, why);
          shape2((z0 - (hz ? zh1 : zh0)) * rdz, whz);

          // first stencil point inside the tile for the centre (i) and edge (h) grids
          const int txi = cx - jx0, txh = txi + hx;
          const int tyi = warp, tyh = warp + hy;
          const int tzi = 0, tzh = hz;
          // Yee staggering, pic/engine/velocity.hpp:442-447
          const double* F = ftile;
          double ex = interp27(F + tzi * FSLAB + (tyi * FX + txh) * 6 + 0, wiz, wiy, whx) * qmdt;
          double ey = interp27(F + tzi * FSLAB + (tyh * FX + txi) * 6 + 1, wiz, why, wix) * qmdt;
          double ez = interp27(F + tzh * FSLAB + (tyi * FX + txi) * 6 + 2, whz, wiy, wix) * qmdt;
          double bx = interp27(F + tzh * FSLAB + (tyh * FX + txi) * 6 + 3, whz, why, wix) * qmdt;
          double by = interp27(F + tzh * FSLAB + (tyi * FX + txh) * 6 + 4, whz, wiy, whx) * qmdt;
          double bz = interp27(F + tzi * FSLAB + (tyh * FX + txh) * 6 + 5, wiz, why, whx) * qmdt;

          push_momentum<Pusher>(ux, uy, uz, ex, ey, ez, bx, by, bz, rc.cc);
          x1 = x0;
          y1 = y0;
          z1 = z0;
          push_position(x1, y1, z1, ux, uy, uz, rc.rc, delt);
          double* xo = PERM ? d.xv : d.xu; // i is the SORTED slot: in place, or the other buffer
          xo[0 * d.pcap + i] = x1;
          xo[1 * d.pcap + i] = y1;
          xo[2 * d.pcap + i] = z1;
          xo[3 * d.pcap + i] = ux;
          xo[4 * d.pcap + i] = uy;
          xo[5 * d.pcap + i] = uz;
          if (PERM)
            xo[6 * d.pcap + i] = pfid;
        } else {
          x0 = d.xv[0 * d.pcap + i];
          y0 = d.xv[1 * d.pcap + i];
          z0 = d.xv[2 * d.pcap + i];
          x1 = d.xu[0 * d.pcap + i];
          y1 = d.xu[1 * d.pcap + i];
          z1 = d.xu[2 * d.pcap + i];
          cx = digitize(x0, xmin, rdx);
          shape2((x0 - (xigrid + (double)cx * dx)) * rdx, s0x);
          shape2((y0 - yig) * rdy, s0y);
          shape2((z0 - zig) * rdz, s0z);
        }

        // new cell: XtensorParticle::count (nix/xtensor_particle.hpp:324-357) and the "after"
        // weights of the Esirkepov scheme share the digitisation (even order: same cell origin)
        const int ix1 = digitize(x1, xmin, rdx);
        const int iy1 = digitize(y1, ymin, rdy);
        const int iz1 = digitize(z1, zmin, rdz);
        if (FUSED) {
          int key = iz1 * g.fsz + iy1 * g.fsy + ix1;
          key     = (x1 < xmin || x1 >= xmax) ? g.Ng : key;
          key     = (y1 < ymin || y1 >= ymax) ? g.Ng : key;
          key     = (z1 < zmin || z1 >= zmax) ? g.Ng : key;
          d.gindex[i] = key;
          atomicAdd(d.pcount + (int64_t)seg * (g.Ng + 1) + key, 1);
          if (key == g.Ng)
            note_leaver(d, seg, i);
        }

        double s1x[3], s1y[3], s1z[3];
        shape2((x1 - (xigrid + (double)ix1 * dx)) * rdx, s1x);
        shape2((y1 - (yigrid + (double)iy1 * dy)) * rdy, s1y);
        shape2((z1 - (zigrid + (double)iz1 * dz)) * rdz, s1z);
        const int shx = ix1 - cx, shy = iy1 - jy, shz = iz1 - jz;
        const int jx  = cx - jx0;
        if (abs(shx) <= 1 && abs(shy) <= 1 && abs(shz) <= 1 && jx >= 0 && jx < RX) {
          const AxisFactors fx = window_factors(s0x, s1x, shx);
          const AxisFactors fy = window_factors(s0y, s1y, shy);
          const AxisFactors fz = window_factors(s0z, s1z, shz);
          stage_particle(ws->stg + lane * REC, fx, fy, fz, q, rc.ddt[2], rc.ddt[1], rc.ddt[0]);
          inf = make_info(jx, fx.w, fy.w, fz.w);
        } else {
          defer_far_mover(d, chunk, q, x0, y0, z0, x1, y1, z1);
        }
      }
      ws->info[lane] = inf;
      if (PERM) {
        cp_async_commit_wait();
        __syncwarp();
      }
      if (FUSED && xbase + lane < xend) {
        const int64_t i = PERM ? xoff + ws->pbuf[lane] : xoff + xbase + lane;
        if (PERM)
          pfid = d.xu[6 * d.pcap + i];
        pfx  = d.xu[0 * d.pcap + i];
        pfy  = d.xu[1 * d.pcap + i];
        pfz  = d.xu[2 * d.pcap + i];
        pfux = d.xu[3 * d.pcap + i];
        pfuy = d.xu[4 * d.pcap + i];
        pfuz = d.xu[5 * d.pcap + i];
      }
      __syncwarp();

      // ---------------- phase 2: one staged particle per half-warp ----------------
      // Which passes can take the register-only fast path is decided here once per batch, without
      // shared-memory loads or votes inside the loop: particle j continues the run of its half-warp
      // iff it has the majority window and the same info word as its predecessor j-2 (for j < 2: as
      // the run carried over from the previous batch).  Bit j of `chg` is set otherwise.
      unsigned chg;
      {
        int pred = __shfl_up_sync(FULL, inf, 2);
        const int carried = __shfl_sync(FULL, curinfo, (lane & 1) << 4);
        if (lane < 2)
          pred = carried;
        const bool runs_on = ((inf >> 8) & 0xf) == 0xf && inf == pred;
        chg                = __ballot_sync(FULL, !runs_on);
      }
      const int npass = (n + 1) >> 1;
      for (int k = 0; k < npass; k++) {
        const int     j   = 2 * k + half;
        const double* rec = ws->stg + j * REC;

        // common case: both particles belong to the cells already being accumulated
        if (((chg >> (2 * k)) & 3u) == 0u) {
          accumulate(acc, rec, a, b);
          continue;
        }
        const int pinf = ws->info[j];

        const bool valid = (pinf >> 11) & 1;
        const int  jx    = pinf & 0xff;
        const int  wx = (pinf >> 8) & 1, wy = (pinf >> 9) & 1, wz = (pinf >> 10) & 1;
        const bool major = valid && (wx & wy & wz);
        const bool minor = valid && !major;

        const bool newcell = major && curinfo != -1 && curinfo != pinf;
        if (__any_sync(FULL, newcell)) {
#pragma unroll
          for (int hh = 0; hh < 2; hh++) {
            if (half == hh && newcell)
              flush(ws->tile, acc, a, b, curinfo & 0xff, 1, 1, 1);
            __syncwarp();
          }
          if (newcell)
            acc.clear();
        }
        if (major) {
          curinfo = pinf;
          accumulate(acc, rec, a, b);
        }
        if (__any_sync(FULL, minor)) {
#pragma unroll
          for (int hh = 0; hh < 2; hh++) {
            if (half == hh && minor)
              deposit_direct(ws->tile, rec, a, b, jx, wx, wy, wz);
            __syncwarp();
          }
        }
      }
      __syncwarp();
    }

    // end of this species' particles in the segment
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      if (half == hh && curinfo != -1)
        flush(ws->tile, acc, a, b, curinfo & 0xff, 1, 1, 1);
      __syncwarp();
    }
    primed = FUSED && pe > pb; // the last batch requested the first one of the next species
    off    = noff;
    pb     = npb;
    pe     = npe;
  }

  // warp tile -> global current: one fp64 reduction per non-zero tile value
  const int gz0 = jz + g.Lb[0] - 2, gy0 = jy + g.Lb[1] - 2, gx0 = jx0 + g.Lb[2] - 2;
#pragma unroll
  for (int comp = 0; comp < 4; comp++) {
    const double* tc = ws->tile + tile_base(comp);
    for (int idx = lane; idx < 25 * XS; idx += 32) {
      const int    tz = idx / (5 * XS);
      const int    r2 = idx - tz * (5 * XS);
      const int    ty = r2 / XS;
      const int    tx = r2 - ty * XS;
      const double v  = tc[tz * tile_sz(comp) + ty * SYT + tx];
      if (v != 0.0)
        atomicAdd(uj + ((int64_t)((gz0 + tz) * My + (gy0 + ty)) * Mx + gx0 + tx) * 4 + comp, v);
    }
  }
}

constexpr int MMA_BLOCKS_PER_SM = 3;

// accumulators of a run (cell + window) -> global uj; info word as in rowdeposit.cuh
template <int ROUND>
__device__ __forceinline__ void flush_run(double* __restrict__ uj, int My, int Mx, const rowmma::Frag& acc,
                                          int fg, int fk, int info, int oz0, int oy0, int ox0)
{
  const int jx = info & 0xff;
  const int wx = (info >> 8) & 1, wy = (info >> 9) & 1, wz = (info >> 10) & 1;
  rowmma::flush_global<ROUND>(uj, My, Mx, acc, fg, fk, oz0 + wz, oy0 + wy, ox0 + jx + wx);
}

// one staging round of a batch: the staged particles, four at a time, into the accumulators
template <int ROUND>
__device__ __forceinline__ void mma_round(rowmma::Frag& acc, int& curinfo, const rowmma::WarpSmem* ws,
                                          int ngroup, int fg, int fk, double* __restrict__ uj, int My,
                                          int Mx, int oz0, int oy0, int ox0)
{
  const unsigned FULL = 0xffffffffu;
  for (int kg = 0; kg < ngroup; kg++) {
    const int my = ws->info[kg * 4 + fk];
    // common case: the four particles continue the run being accumulated
    if (__all_sync(FULL, my == curinfo)) {
      rowmma::mma_group(acc, ws->stg, kg, fg, fk, true);
      continue;
    }
    bool pending = (my >> 11) & 1;
    while (__any_sync(FULL, pending)) {
      const unsigned m   = __ballot_sync(FULL, pending);
      const int      cur = __shfl_sync(FULL, my, __ffs(m) - 1);
      const bool     sel = pending && my == cur;
      if (((cur >> 8) & 7) == 7) {
        // majority window: a new run starts when the cell changes
        if (cur != curinfo) {
          if (curinfo != -1)
            flush_run<ROUND>(uj, My, Mx, acc, fg, fk, curinfo, oz0, oy0, ox0);
          acc.clear();
          curinfo = cur;
        }
        rowmma::mma_group(acc, ws->stg, kg, fg, fk, sel);
      } else {
        // window shifted down in some axis (particle moved to the lower cell): on its own
        rowmma::Frag tmp;
        tmp.clear();
        rowmma::mma_group(tmp, ws->stg, kg, fg, fk, sel);
        flush_run<ROUND>(uj, My, Mx, tmp, fg, fk, cur, oz0, oy0, ox0);
      }
      pending = pending && !sel;
    }
  }
}

template <bool FUSED, int Pusher, int Interp>
__global__ void __launch_bounds__(THREADS, MMA_BLOCKS_PER_SM)
row_mma_kernel(Geom g, DevPtrs d, RowConst rc, int c0, int cn, double delt)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double*   ftile = reinterpret_cast<double*>(smem_raw);
  rowmma::WarpSmem* wsm = reinterpret_cast<rowmma::WarpSmem*>(smem_raw + sizeof(double) * FTILE);

  const int      lane = threadIdx.x & 31;
  const int      warp = threadIdx.x >> 5;
  const int      fg   = lane >> 2; // MMA fragment row / column group
  const int      fk   = lane & 3;  // MMA contraction index (particle inside a group of 4)
  const unsigned FULL = 0xffffffffu;
  rowmma::WarpSmem* ws = wsm + warp;

  // block -> (chunk, z, y group, x segment)
  const int nsegx = g.dims[2] / RX;
  const int nygrp = g.dims[1] / WARPS;
  int       r     = blockIdx.x;
  const int lc    = r / (g.dims[0] * nygrp * nsegx);
  r -= lc * g.dims[0] * nygrp * nsegx;
  const int jz = r / (nygrp * nsegx);
  r -= jz * nygrp * nsegx;
  const int jy0   = (r / nsegx) * WARPS;
  const int jx0   = (r - (r / nsegx) * nsegx) * RX;
  const int jy    = jy0 + warp;
  const int chunk = c0 + lc;

  const double* lim = d.clim + chunk * 6;
  double*       uj  = d.uj + (int64_t)chunk * g.Ng * 4;
  const int     My = g.M[1], Mx = g.M[2];

  // ---- stage the field tile (global layout [z][y][x][6], 16-byte copies) ----
  if (FUSED) {
    const double* uf = d.uf + (int64_t)chunk * g.Ng * 6;
    const int     gz = jz + g.Lb[0] - 1, gy = jy0 + g.Lb[1] - 1, gx = jx0 + g.Lb[2] - 1;
    for (int e = threadIdx.x; e < FZ * FY * (FROW / 2); e += THREADS) {
      const int row = e / (FROW / 2);
      const int col = e - row * (FROW / 2);
      const int tz  = row / FY;
      const int ty  = row - tz * FY;
      const double2 v = __ldg(reinterpret_cast<const double2*>(
                                  uf + ((int64_t)((gz + tz) * My + (gy + ty)) * Mx + gx) * 6) + col);
      reinterpret_cast<double2*>(ftile + tz * FSLAB + ty * FROW)[col] = v;
    }
  }
  if (FUSED)
    __syncthreads();
  else
    __syncwarp();

  const double xmin = lim[4], ymin = lim[2], zmin = lim[0];
  const double xmax = lim[5], ymax = lim[3], zmax = lim[1]; // in registers: the key tests run per particle
  const double rdx = rc.rd[2], rdy = rc.rd[1], rdz = rc.rd[0];
  const double dx = rc.del[2], dy = rc.del[1], dz = rc.del[0];
  // cell-centre ("integer") and cell-edge ("half") grid points of this row, pic/engine/velocity.hpp:304-315
  const double yig = ymin + 0.5 * dy + (double)jy * dy;
  const double zig = zmin + 0.5 * dz + (double)jz * dz;
  const double yh0 = ymin + (double)jy * dy, yh1 = ymin + (double)(jy + 1) * dy;
  const double zh0 = zmin + (double)jz * dz, zh1 = zmin + (double)(jz + 1) * dz;
  const double xigrid = xmin + 0.5 * dx, yigrid = ymin + 0.5 * dy, zigrid = zmin + 0.5 * dz;
  const int    key0 = jz * g.fsz + jy * g.fsy + jx0;
  // global index of stencil slot 0 (two cells left of / below the cell) of this row's first cell
  const int    oz0 = jz + g.Lb[0] - 2, oy0 = jy + g.Lb[1] - 2, ox0 = jx0 + g.Lb[2] - 2;

  for (int is = 0; is < g.Ns; is++) {
    const int     seg = chunk * g.Ns + is;
    const int64_t off = d.seg_off[seg];
    const int*    pix = d.pindex + (int64_t)seg * (g.Ng + 1);
    const int     pb = pix[key0], pe = pix[key0 + RX];
    const double  q    = d.qm[2 * is];
    const double  qmdt = 0.5 * q / d.qm[2 * is + 1] * delt;

    rowmma::Frag acc0, acc1; // round 0: rho, Jx; round 1: Jy, Jz
    acc0.clear();
    acc1.clear();
    int curinfo0 = -1, curinfo1 = -1; // info word of the cell each accumulator set belongs to

    // the six phase-space components of the NEXT batch are requested before phase 2 of the current
    // one, so their HBM latency is hidden behind the accumulation loop
    double pfx = 0, pfy = 0, pfz = 0, pfux = 0, pfuy = 0, pfuz = 0;
    if (FUSED && pb + lane < pe) {
      const int64_t i = off + pb + lane;
      pfx  = d.xu[0 * d.pcap + i];
      pfy  = d.xu[1 * d.pcap + i];
      pfz  = d.xu[2 * d.pcap + i];
      pfux = d.xu[3 * d.pcap + i];
      pfuy = d.xu[4 * d.pcap + i];
      pfuz = d.xu[5 * d.pcap + i];
    }

    for (int base = pb; base < pe; base += 32) {
      const int n = min(32, pe - base);

      // ---------------- phase 1: one particle per lane ----------------
      int             inf = 0;
      rowmma::Factors fac;
      if (lane < n) {
        const int64_t i = off + base + lane;
        double        x0, y0, z0, x1, y1, z1;
        double        s0x[3], s0y[3], s0z[3];
        int           cx; // old cell in x, relative to the chunk
        if (FUSED) {
          x0        = pfx;
          y0        = pfy;
          z0        = pfz;
          double ux = pfux;
          double uy = pfuy;
          double uz = pfuz;

          // weights on the centre grid (MC or WT) and on the edge grid (MC); the particle is in
          // row (jz, jy) by construction of the sort, only its x cell has to be found
          double wix[3], whx[3], wiy[3], why[3], wiz[3], whz[3];
          cx = digitize(x0, xmin, rdx);
          const double cxf = (double)cx;
          const double dix = (x0 - (xigrid + cxf * dx)) * rdx;
          const double diy = (y0 - yig) * rdy;
          const double diz = (z0 - zig) * rdz;
          shape2(dix, s0x);
          shape2(diy, s0y);
          shape2(diz, s0z);
          if (Interp == PICNIX_INTERP_MC) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
              wix[k] = s0x[k];
              wiy[k] = s0y[k];
              wiz[k] = s0z[k];
            }
          } else {
            shape_wt<2>(x0, xigrid + cxf * dx, rdx, rc.cfl[2], 1 / rc.cfl[2], wix);
            shape_wt<2>(y0, yig, rdy, rc.cfl[1], 1 / rc.cfl[1], wiy);
            shape_wt<2>(z0, zig, rdz, rc.cfl[0], 1 / rc.cfl[0], wiz);
          }
          // nearest cell edge: the one to the right when the particle sits right of the centre
          const int hx = dix >= 0.0, hy = diy >= 0.0, hz = diz >= 0.0;
          shape2((x0 - (xmin + (cxf + (double)hx) * dx)) * rdx, whx);
          shape2((y0 - (hy ? yh1 : yh0)) * rdy, why);
          shape2((z0 - (hz ? zh1 : zh0)) * rdz, whz);

          // first stencil point inside the tile for the centre (i) and edge (h) grids
          const int txi = cx - jx0, txh = txi + hx;
          const int tyi = warp, tyh = warp + hy;
          const int tzi = 0, tzh = hz;
          // Yee staggering, pic/engine/velocity.hpp:442-447
          const double* F = ftile;
          double ex = interp27(F + tzi * FSLAB + (tyi * FX + txh) * 6 + 0, wiz, wiy, whx) * qmdt;
          double ey = interp27(F + tzi * FSLAB + (tyh * FX + txi) * 6 + 1, wiz, why, wix) * qmdt;
          double ez = interp27(F + tzh * FSLAB + (tyi * FX + txi) * 6 + 2, whz, wiy, wix) * qmdt;
          double bx = interp27(F + tzh * FSLAB + (tyh * FX + txi) * 6 + 3, whz, why, wix) * qmdt;
          double by = interp27(F + tzh * FSLAB + (tyi * FX + txh) * 6 + 4, whz, wiy, whx) * qmdt;
          double bz = interp27(F + tzi * FSLAB + (tyh * FX + txh) * 6 + 5, wiz, why, whx) * qmdt;

          push_momentum<Pusher>(ux, uy, uz, ex, ey, ez, bx, by, bz, rc.cc);
          x1 = x0;
          y1 = y0;
          z1 = z0;
          push_position(x1, y1, z1, ux, uy, uz, rc.rc, delt);
          d.xu[0 * d.pcap + i] = x1;
          d.xu[1 * d.pcap + i] = y1;
          d.xu[2 * d.pcap + i] = z1;
          d.xu[3 * d.pcap + i] = ux;
          d.xu[4 * d.pcap + i] = uy;
          d.xu[5 * d.pcap + i] = uz;
        } else {
          x0 = d.xv[0 * d.pcap + i];
          y0 = d.xv[1 * d.pcap + i];
          z0 = d.xv[2 * d.pcap + i];
          x1 = d.xu[0 * d.pcap + i];
          y1 = d.xu[1 * d.pcap + i];
          z1 = d.xu[2 * d.pcap + i];
          cx = digitize(x0, xmin, rdx);
          shape2((x0 - (xigrid + (double)cx * dx)) * rdx, s0x);
          shape2((y0 - yig) * rdy, s0y);
          shape2((z0 - zig) * rdz, s0z);
        }

        // new cell: XtensorParticle::count (nix/xtensor_particle.hpp:324-357) and the "after"
        // weights of the Esirkepov scheme share the digitisation (even order: same cell origin)
        const int ix1 = digitize(x1, xmin, rdx);
        const int iy1 = digitize(y1, ymin, rdy);
        const int iz1 = digitize(z1, zmin, rdz);
        if (FUSED) {
          int key = iz1 * g.fsz + iy1 * g.fsy + ix1;
          key     = (x1 < xmin || x1 >= xmax) ? g.Ng : key;
          key     = (y1 < ymin || y1 >= ymax) ? g.Ng : key;
          key     = (z1 < zmin || z1 >= zmax) ? g.Ng : key;
          d.gindex[i] = key;
          atomicAdd(d.pcount + (int64_t)seg * (g.Ng + 1) + key, 1);
          if (key == g.Ng)
            note_leaver(d, seg, i);
        }

        double s1x[3], s1y[3], s1z[3];
        shape2((x1 - (xigrid + (double)ix1 * dx)) * rdx, s1x);
        shape2((y1 - (yigrid + (double)iy1 * dy)) * rdy, s1y);
        shape2((z1 - (zigrid + (double)iz1 * dz)) * rdz, s1z);
        const int shx = ix1 - cx, shy = iy1 - jy, shz = iz1 - jz;
        const int jx  = cx - jx0;
        if (abs(shx) <= 1 && abs(shy) <= 1 && abs(shz) <= 1 && jx >= 0 && jx < RX) {
          const AxisFactors fx = window_factors(s0x, s1x, shx);
          const AxisFactors fy = window_factors(s0y, s1y, shy);
          const AxisFactors fz = window_factors(s0z, s1z, shz);
          fac = rowmma::make_factors(fx, fy, fz, q, rc.ddt[2], rc.ddt[1], rc.ddt[0]);
          inf = make_info(jx, fx.w, fy.w, fz.w);
        } else {
          defer_far_mover(d, chunk, q, x0, y0, z0, x1, y1, z1);
        }
      }
      ws->info[lane] = inf;
      if (FUSED && base + 32 + lane < pe) {
        const int64_t i = off + base + 32 + lane;
        pfx  = d.xu[0 * d.pcap + i];
        pfy  = d.xu[1 * d.pcap + i];
        pfz  = d.xu[2 * d.pcap + i];
        pfux = d.xu[3 * d.pcap + i];
        pfuy = d.xu[4 * d.pcap + i];
        pfuz = d.xu[5 * d.pcap + i];
      }
      __syncwarp();

      // ---------------- phase 2: groups of 4 staged particles through the FP64 MMA ----------------
      const int ngroup = (n + 3) >> 2;
      if (inf)
        rowmma::stage_round<0>(ws->stg, lane, fac);
      __syncwarp();
      mma_round<0>(acc0, curinfo0, ws, ngroup, fg, fk, uj, My, Mx, oz0, oy0, ox0);
      __syncwarp();
      if (inf)
        rowmma::stage_round<1>(ws->stg, lane, fac);
      __syncwarp();
      mma_round<1>(acc1, curinfo1, ws, ngroup, fg, fk, uj, My, Mx, oz0, oy0, ox0);
      __syncwarp();
    }

    // end of this species' particles in the segment
    if (curinfo0 != -1)
      flush_run<0>(uj, My, Mx, acc0, fg, fk, curinfo0, oz0, oy0, ox0);
    if (curinfo1 != -1)
      flush_run<1>(uj, My, Mx, acc1, fg, fk, curinfo1, oz0, oy0, ox0);
  }
}

template <bool FUSED>
int launch_row_kernel(picnix_arena* a, int c0, int cn, double delt)
{
  const Geom& g      = a->g;
  const int   blocks = g.dims[0] * (g.dims[1] / WARPS) * (g.dims[2] / RX) * cn;
  const int   key    = FUSED ? a->cfg.pusher * 2 + a->cfg.interp : 0;
  // a pending lazy sort is consumed by the fused kernel itself when it covers the whole arena;
  // everything else (partial ranges, deposit only, the MMA experiment) wants physically ordered arrays
  const bool  perm   = FUSED && a->perm_pending && c0 == 0 && cn == g.nchunk && !a->deposit_mma;
  if (!perm) {
    int status = materialize_sort(a);
    if (status != PICNIX_OK)
      return status;
  }

  RowConst rc;
  for (int i = 0; i < 3; i++) {
    rc.rd[i]  = 1 / g.del[i];
    rc.del[i] = g.del[i];
    rc.ddt[i] = g.del[i] / delt;
    rc.cfl[i] = g.cc * delt / g.del[i];
  }
  rc.cc   = g.cc;
  rc.rc   = 1 / g.cc;
  rc.delt = delt;
  PICNIX_CUDA(a, cudaMemsetAsync(a->d.far_count, 0, sizeof(int), a->stream));

  // the FP64-MMA deposit is an experiment kept for the record (slower, see rowmma.cuh); it is only
  // instantiated for the benchmark configuration (fused, Boris, MC)
#define PICNIX_ROW_LAUNCH(P, I)                                                                    \
  if (a->deposit_mma && FUSED && P == PICNIX_PUSHER_BORIS && I == PICNIX_INTERP_MC) {              \
    auto kern = row_mma_kernel<true, PICNIX_PUSHER_BORIS, PICNIX_INTERP_MC>;                       \
    PICNIX_CUDA(a, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                        (int)rowmma::SMEM_BYTES));                                 \
    kern<<<blocks, THREADS, rowmma::SMEM_BYTES, a->stream>>>(g, a->d, rc, c0, cn, delt);           \
  } else if (perm) {                                                                               \
    auto kern = row_kernel<FUSED, P, I, FUSED>;                                                    \
    PICNIX_CUDA(a, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                        (int)SMEM_BYTES));                                         \
    kern<<<blocks, THREADS, SMEM_BYTES, a->stream>>>(g, a->d, rc, c0, cn, delt);                   \
  } else {                                                                                         \
    auto kern = row_kernel<FUSED, P, I, false>;                                                    \
    PICNIX_CUDA(a, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                        (int)SMEM_BYTES));                                         \
    kern<<<blocks, THREADS, SMEM_BYTES, a->stream>>>(g, a->d, rc, c0, cn, delt);                   \
  }
  if constexpr (!FUSED) {
    // deposit only: pusher and interpolation do not enter, one instantiation serves all
    PICNIX_ROW_LAUNCH(PICNIX_PUSHER_BORIS, PICNIX_INTERP_MC);
  } else {
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
  }
#undef PICNIX_ROW_LAUNCH
  far_kernel<<<64, 128, 0, a->stream>>>(g, a->d, delt);
  a->kernel_launches += 2;
  if (perm) {
    // the kernel wrote the pushed particles in sorted order into xv
    std::swap(a->d.xu, a->d.xv);
    a->perm_pending = false;
  }
  return check_cuda(a, cudaGetLastError(), "row_kernel");
}

} // namespace

int launch_row_kernel_v1(picnix_arena* a, int c0, int cn, double delt, bool fused)
{
  return fused ? launch_row_kernel<true>(a, c0, cn, delt) : launch_row_kernel<false>(a, c0, cn, delt);
}

// v1 geometry: rows that split into RX-cell segments and WARPS-row groups
bool row_v1_geometry(const picnix_arena* a)
{
  const Geom& g = a->g;
  return g.dimension == 3 && g.order == 2 && (g.dims[2] % rowdep::RX) == 0 && (g.dims[1] % rowdep::WARPS) == 0;
}

} // namespace picnix
