// -*- C++ -*-
// Row-owner kernel for 2-D runs (x, y; z ignorable), order 2: the two-dimensional variant of rowpush.cu,
// see rowtile2d.cuh for the scheme.  One block per (chunk, group of WARPS rows in y, x-segment of RX
// cells), one warp per row; FUSED / PERM as in rowpush.cu.
#include "rowtile2d.cuh"

namespace picnix
{

namespace
{

using namespace rowtile2d;

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
    esirkepov_deposit<2, 2>(g, lim, r[6], delt, r[0], r[1], r[2], r[3], r[4], r[5], bz, by, bx, add);
  }
}

// One slot of the merged particle stream of a row segment: which particle, if any.
//   idx  index inside the (chunk, species) segment, -1 for the idle slot that pads a cell to ALIGN
//   sc   species | cell << 8  (cell relative to the segment)
struct Slot {
  int idx, sc;
};

// slot t of the stream; k is the lane's cursor into the entry table (slots are asked for in
// ascending order, an entry is about one batch long: the loop runs once or twice)
__device__ __forceinline__ Slot stream_slot(const WarpSmem* ws, int t, int nent, int& k)
{
  while (k < nent && t >= ws->ent[k + 1].x)
    k++;
  Slot s;
  s.idx = -1;
  s.sc  = 0;
  if (k < nent) {
    const int4 e = ws->ent[k];
    const int  r = t - e.x;
    if (r < e.z) {
      s.idx = e.y + r;
      s.sc  = e.w;
    }
  }
  return s;
}

// PERM (fused only): a lazy sort is pending -- sorted slot j of a segment still sits in slot perm[j] of
// xu; the kernel reads through the permutation and writes the pushed particle (all seven components)
// to slot j of xv, so the reordering costs no pass of its own (the host swaps xu/xv afterwards).
template <bool FUSED, int Pusher, int Interp, bool PERM>
__global__ void __launch_bounds__(THREADS, 3)
row_push2d_kernel(Geom g, DevPtrs d, RowConst rc, int c0, int cn, double delt)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double*    ftile = reinterpret_cast<double*>(smem_raw);
  BlockSmem* bs    = reinterpret_cast<BlockSmem*>(smem_raw + sizeof(double) * FTILE);
  WarpSmem*  wsm   = reinterpret_cast<WarpSmem*>(smem_raw + sizeof(double) * FTILE + sizeof(BlockSmem));

  const int      lane = threadIdx.x & 31;
  const int      warp = threadIdx.x >> 5;
  const int      half = lane >> 4;
  const unsigned FULL = 0xffffffffu;
  WarpSmem*      ws   = wsm + warp;
  const LaneMap  lm   = lane_map(lane);
  const int      Ns   = g.Ns;

  // block -> (chunk, y group, x segment)
  const int nsegx = g.dims[2] / RX;
  const int nygrp = g.dims[1] / WARPS;
  int       r     = blockIdx.x;
  const int lc    = r / (nygrp * nsegx);
  r -= lc * nygrp * nsegx;
  const int jy0   = (r / nsegx) * WARPS;
  const int jx0   = (r - (r / nsegx) * nsegx) * RX;
  const int jy    = jy0 + warp;
  const int chunk = c0 + lc;

  const double* lim = d.clim + chunk * 6;
  double*       uj  = d.uj + (int64_t)chunk * g.Ng * 4;
  const int     My = g.M[1], Mx = g.M[2];

  // ---- the field tile starts travelling (the one z plane, layout [y][x][6], 16-byte asynchronous copies) ----
  if (FUSED) {
    const double* uf = d.uf + (int64_t)chunk * g.Ng * 6;
    const int     gz = g.Lb[0], gy = jy0 + g.Lb[1] - 1, gx = jx0 + g.Lb[2] - 1;
    for (int e = threadIdx.x; e < FY * (FROW / 2); e += THREADS) {
      const int ty  = e / (FROW / 2);
      const int col = e - ty * (FROW / 2);
      cp_async_16(reinterpret_cast<double2*>(ftile + ty * FROW) + col,
                  reinterpret_cast<const double2*>(uf + ((int64_t)(gz * My + (gy + ty)) * Mx + gx) * 6) + col);
    }
  }
  if (threadIdx.x < Ns) {
    const int    is = threadIdx.x;
    const double q  = d.qm[2 * is];
    bs->q[is]       = q;
    bs->qmdt[is]    = 0.5 * q / d.qm[2 * is + 1] * delt;
    bs->off[is]     = d.seg_off[chunk * Ns + is];
  }

  // ---- the stream of this warp's row segment: lane (cell c, species is) = c * Ns + is builds its entry;
  // cells are padded to a multiple of ALIGN slots ----
  const int key0 = jy * g.fsy + jx0;
  const int nent = RX * Ns;
  {
    const int c  = lane / Ns;
    const int is = lane - c * Ns;
    int       b = 0, n = 0;
    if (lane < nent) {
      const int* pix = d.pindex + (int64_t)(chunk * Ns + is) * (g.Ng + 1) + key0 + c;
      b              = pix[0];
      n              = pix[1] - b;
    }
    // exclusive prefix of the counts over the lanes
    int incl = n;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
      const int t = __shfl_up_sync(FULL, incl, dd);
      if (lane >= dd)
        incl += t;
    }
    // padding accumulated before cell c: every earlier cell rounds its total up to ALIGN
    const int cellend = __shfl_sync(FULL, incl, min(c * Ns + Ns - 1, 31)); // slots of cells 0..c
    const int celltot = cellend - __shfl_sync(FULL, incl - n, min(c * Ns, 31));
    int       pad     = (lane < nent && is == Ns - 1) ? ((celltot + ALIGN - 1) & ~(ALIGN - 1)) - celltot : 0;
    int       pincl   = pad;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
      const int t = __shfl_up_sync(FULL, pincl, dd);
      if (lane >= dd)
        pincl += t;
    }
    const int start = (incl - n) + (pincl - pad); // unpadded start + padding of the cells before
    if (lane < nent)
      ws->ent[lane] = make_int4(start, b, n, is | (c << 8));
    if (lane == nent - 1)
      ws->ent[nent] = make_int4(start + n + pad, 0, 0, 0);
  }
  for (int i = lane; i < TILE; i += 32)
    ws->tile[i] = 0.0;
  for (int i = lane; i < REC + 2; i += 32)
    ws->zero[i] = 0.0;
  __syncthreads(); // stream tables and species constants visible (the field tile is still in flight)

  const int total = ws->ent[nent].x;

  // ---- first batch: its phase space travels global -> shared behind the field tile; so do the
  // permutation entries of the second batch (everything asynchronous, nothing held in registers) ----
  int  kent = 0;
  Slot cur  = stream_slot(ws, lane, nent, kent);
  Slot nxt  = stream_slot(ws, 32 + lane, nent, kent);
  if (FUSED) {
    if (cur.idx >= 0) {
      const int64_t off = bs->off[cur.sc & 0xff];
      const int64_t i   = PERM ? off + d.perm[off + cur.idx] : off + cur.idx;
#pragma unroll
      for (int k = 0; k < (PERM ? 7 : 6); k++)
        cp_async_f64(&ws->pfb[k][lane], d.xu + k * d.pcap + i);
    }
    if (PERM && nxt.idx >= 0)
      cp_async_i32(ws->pbuf + lane, d.perm + bs->off[nxt.sc & 0xff] + nxt.idx);
    cp_async_commit_wait();
    __syncthreads();
  }

  const double rdx = rc.rd[2], rdy = rc.rd[1];
  const double dx = rc.del[2], dy = rc.del[1];
  // chunk limits and grid points of the row live in shared memory and are re-read at every use (see
  // rowpush.cu): cell-centre ("integer") and cell-edge ("half") points, pic/engine/velocity.hpp:304-315
  if (lane == 0) {
    const double xmin0 = lim[4], ymin0 = lim[2];
    ws->rowc[0] = xmin0;
    ws->rowc[1] = ymin0;
    ws->rowc[2] = lim[5];
    ws->rowc[3] = lim[3];
    ws->rowc[4] = ymin0 + 0.5 * dy + (double)jy * dy; // yig
    ws->rowc[5] = ymin0 + (double)jy * dy;            // yh0
    ws->rowc[6] = ymin0 + (double)(jy + 1) * dy;      // yh1
    ws->rowc[7] = xmin0 + 0.5 * dx;                   // xigrid
    ws->rowc[8] = ymin0 + 0.5 * dy;                   // yigrid
  }
  __syncwarp();
  const volatile double* rowc = ws->rowc;
#define xmin rowc[0]
#define ymin rowc[1]
#define xmax rowc[2]
#define ymax rowc[3]
#define yig rowc[4]
#define yh0 rowc[5]
#define yh1 rowc[6]
#define xigrid rowc[7]
#define yigrid rowc[8]

  Acc acc;
  acc.clear();
  int curinfo = -1; // info word of the cell the accumulators belong to (-1: none)

  for (int base = 0; base < total; base += 32) {
    // this batch's phase space has landed in shared memory (and the permutation entries of the next one)
    double pfx = 0, pfy = 0, pfz = 0, pfux = 0, pfuy = 0, pfuz = 0, pfid = 0;
    if (FUSED && cur.idx >= 0) {
      pfx  = ws->pfb[0][lane];
      pfy  = ws->pfb[1][lane];
      pfz  = ws->pfb[2][lane];
      pfux = ws->pfb[3][lane];
      pfuy = ws->pfb[4][lane];
      pfuz = ws->pfb[5][lane];
      if (PERM)
        pfid = ws->pfb[6][lane];
    }
    // the next batch starts travelling now and has phases 1 and 2 of this one to arrive; the permutation
    // entries are requested two batches ahead
    const Slot nn = stream_slot(ws, base + 64 + lane, nent, kent);
    if (FUSED) {
      if (nxt.idx >= 0) {
        const int64_t off = bs->off[nxt.sc & 0xff];
        const int64_t i   = PERM ? off + ws->pbuf[lane] : off + nxt.idx;
#pragma unroll
        for (int k = 0; k < (PERM ? 7 : 6); k++)
          cp_async_f64(&ws->pfb[k][lane], d.xu + k * d.pcap + i);
      }
      if (PERM && nn.idx >= 0)
        cp_async_i32(ws->pbuf + lane, d.perm + bs->off[nn.sc & 0xff] + nn.idx);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }

    // ---------------- phase 1: one particle per lane ----------------
    int inf = 0;
    if (cur.idx >= 0) {
      const int     is = cur.sc & 0xff;
      const int     jx = cur.sc >> 8;  // old cell in x relative to the segment: given by the sort
      const int     cx = jx0 + jx;
      const int64_t i  = bs->off[is] + cur.idx;
      const double  q  = bs->q[is];
      double        x0, y0, z0, x1, y1, z1;
      const double  cxf = (double)cx;
      if (FUSED) {
        x0        = pfx;
        y0        = pfy;
        z0        = pfz;
        double ux = pfux;
        double uy = pfuy;
        double uz = pfuz;

        // weights on the centre grid (MC or WT) and on the edge grid (MC); the particle is in cell
        // (jy, cx) by construction of the sort
        double s0x[3], s0y[3];
        double wix[3], wiy[3], h[3], whx[4], why[4];
        const double dix = (x0 - (xigrid + cxf * dx)) * rdx;
        const double diy = (y0 - yig) * rdy;
        shape2(dix, s0x);
        shape2(diy, s0y);
        if (Interp == PICNIX_INTERP_MC) {
#pragma unroll
          for (int k = 0; k < 3; k++) {
            wix[k] = s0x[k];
            wiy[k] = s0y[k];
          }
        } else {
          shape_wt<2>(x0, xigrid + cxf * dx, rdx, rc.cfl[2], 1 / rc.cfl[2], wix);
          shape_wt<2>(y0, yig, rdy, rc.cfl[1], 1 / rc.cfl[1], wiy);
        }
        // nearest cell edge: the one to the right when the particle sits right of the centre; its
        // three weights go into the cell-anchored 4-slot array
        const bool hx = dix >= 0.0, hy = diy >= 0.0;
        shape2((x0 - (xmin + (cxf + (hx ? 1.0 : 0.0)) * dx)) * rdx, h);
        shift4(h, hx, whx);
        shape2((y0 - (hy ? yh1 : yh0)) * rdy, h);
        shift4(h, hy, why);

        // first stencil point of the cell in the tile; Yee staggering, pic/engine/velocity.hpp:410-415
        const double* F    = ftile + warp * FROW + jx * 6;
        const double  qmdt = bs->qmdt[is];
        double ex = interp_cell<3, 4>(F + 0, wiy, whx) * qmdt;
        double ey = interp_cell<4, 3>(F + 1, why, wix) * qmdt;
        double ez = interp_cell<3, 3>(F + 2, wiy, wix) * qmdt;
        double bx = interp_cell<4, 3>(F + 3, why, wix) * qmdt;
        double by = interp_cell<3, 4>(F + 4, wiy, whx) * qmdt;
        double bz = interp_cell<4, 4>(F + 5, why, whx) * qmdt;

        if (Pusher == PICNIX_PUSHER_BORIS)
          push_boris_fast(ux, uy, uz, ex, ey, ez, bx, by, bz, rc.cc);
        else
          push_momentum<Pusher>(ux, uy, uz, ex, ey, ez, bx, by, bz, rc.cc);
        x1 = x0;
        y1 = y0;
        z1 = z0;
        push_position_fast(x1, y1, z1, ux, uy, uz, rc.rc, delt);
        apply_particle_bc(g, x1, y1, z1, ux, uy, uz);
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
      }

      // new cell: XtensorParticle::count (nix/xtensor_particle.hpp:324-357) and the "after"
      // weights of the Esirkepov scheme share the digitisation (even order: same cell origin)
      const int ix1 = digitize(x1, xmin, rdx);
      const int iy1 = digitize(y1, ymin, rdy);
      if (FUSED) {
        const int seg = chunk * Ns + is;
        int       key = iy1 * g.fsy + ix1;
        key           = (x1 < xmin || x1 >= xmax) ? g.Ng : key;
        key           = (y1 < ymin || y1 >= ymax) ? g.Ng : key;
        d.gindex[i]   = key;
        atomicAdd(d.pcount + (int64_t)seg * (g.Ng + 1) + key, 1);
        if (key == g.Ng)
          note_leaver(d, seg, i);
      }

      double s0x[3], s0y[3], s1x[3], s1y[3];
      shape2((x0 - (xigrid + cxf * dx)) * rdx, s0x);
      shape2((y0 - yig) * rdy, s0y);
      shape2((x1 - (xigrid + (double)ix1 * dx)) * rdx, s1x);
      shape2((y1 - (yigrid + (double)iy1 * dy)) * rdy, s1y);
      const int shx = ix1 - cx, shy = iy1 - jy;
      if (abs(shx) <= 1 && abs(shy) <= 1) {
        const AxisFactors fx = window_factors(s0x, s1x, shx);
        const AxisFactors fy = window_factors(s0y, s1y, shy);
        // the ignorable direction contributes through the velocity (pic/engine/current.hpp:270-274)
        const double qvz = q * ((z1 - z0) / delt);
        stage_particle(ws->stg + lane * REC, fx, fy, q, qvz, rc.ddt[2], rc.ddt[1]);
        inf = make_info(jx, fx.w, fy.w, 1);
      } else {
        defer_far_mover(d, chunk, q, x0, y0, z0, x1, y1, z1);
      }
    }
    ws->info[lane] = inf;

    // ---- the cells of the batch: lanes are in stream order, so the particles of a cell that have the
    // majority window (the common case) form one ascending lane range, interrupted only by the few
    // particles with another window and by the idle slot that pads a cell
    const bool     major = ((inf >> 8) & 0xf) == 0xf;
    const unsigned mm    = __ballot_sync(FULL, major);
    const unsigned om    = __ballot_sync(FULL, inf != 0 && !major);
    unsigned       same  = 0;
    if (major)
      same = __match_any_sync(mm, inf);
    const unsigned leaders = __ballot_sync(FULL, major && (__ffs(same) - 1) == lane);
    __syncwarp();

    // ---------------- phase 2: one staged particle per half-warp ----------------
    // Cell by cell (warp-uniform control): when the cell differs from the one the accumulators belong
    // to, both half-warps add their patches to the tile; then the lane range of the cell is consumed two
    // records per pass, the lower half-warp the first, the upper one the second.  A record that is not
    // a majority-window particle of the cell is replaced by the all-zero record.
    for (unsigned gl = leaders; gl != 0; gl &= gl - 1) {
      const int      L     = __ffs(gl) - 1;
      const int      ginfo = __shfl_sync(FULL, inf, L);
      const unsigned gm    = __shfl_sync(FULL, same, L);
      const int      last  = 31 - __clz(gm);
      if (ginfo != curinfo) {
        if (curinfo != -1) {
          flush(ws->tile, acc, lm, run_index(curinfo), half);
          __syncwarp();
          acc.clear();
        }
        curinfo = ginfo;
      }
      const int cnt = last - L + 1;
      if (__popc(gm) == cnt) {
        // no foreign slot inside the range (the usual case): plain pointer walk
        const double* rec = ws->stg + (L + half) * REC;
#pragma unroll 2
        for (int k = 0; k < (cnt >> 1); k++) {
          accumulate(acc, rec, lm);
          rec += 2 * REC;
        }
        if (cnt & 1)
          accumulate(acc, half == 0 ? rec : ws->zero, lm);
      } else {
#pragma unroll 1
        for (int j = L + half; j <= last + half; j += 2) {
          const double* rec = ((gm >> (j & 31)) & 1u) && j <= last ? ws->stg + j * REC : ws->zero;
          accumulate(acc, rec, lm);
        }
      }
    }
    // the few particles with another window (moved to the lower cell in some direction): straight into
    // the tile, the whole warp on one record (each half-warp two of the four rows of every patch)
    for (unsigned mk = om; mk != 0; mk &= mk - 1) {
      const int j = __ffs(mk) - 1;
      deposit_direct(ws->tile, ws->stg + j * REC, lm, run_index(ws->info[j]), half);
      __syncwarp(); // the next record's window may overlap this one's: other lanes, same tile elements
    }
    if (FUSED)
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    cur = nxt;
    nxt = nn;
  }

  // end of the segment: the accumulators of both half-warps
  if (curinfo != -1)
    flush(ws->tile, acc, lm, run_index(curinfo), half);
  __syncwarp();

  // warp tile -> global current: one fp64 reduction per non-zero tile value; a y line of the tile is one
  // contiguous run of uj (the four components of a point stay together in both)
  const int gz0 = g.Lb[0], gy0 = jy + g.Lb[1] - 2, gx0 = jx0 + g.Lb[2] - 2;
  for (int ty = 0; ty < 5; ty++) {
    const double* src = ws->tile + 4 * (ty * SY);
    double*       dst = uj + ((int64_t)(gz0 * My + (gy0 + ty)) * Mx + gx0) * 4;
    for (int e = lane; e < 4 * XS; e += 32) {
      const double v = src[e];
      if (v != 0.0)
        atomicAdd(dst + e, v);
    }
  }
}

#undef xmin
#undef ymin
#undef xmax
#undef ymax
#undef yig
#undef yh0
#undef yh1
#undef xigrid
#undef yigrid

template <bool FUSED>
int launch_row_kernel(picnix_arena* a, int c0, int cn, double delt)
{
  const Geom& g      = a->g;
  const int   blocks = (g.dims[1] / WARPS) * (g.dims[2] / RX) * cn;
  const int   key    = FUSED ? a->cfg.pusher * 2 + a->cfg.interp : 0;
  // a pending lazy sort is consumed by the fused kernel itself when it covers the whole arena;
  // everything else (partial ranges, deposit only) wants physically ordered arrays
  const bool  perm   = FUSED && a->perm_pending && c0 == 0 && cn == g.nchunk;
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

#define PICNIX_ROW_LAUNCH(P, I)                                                                    \
  if (perm) {                                                                                      \
    auto kern = row_push2d_kernel<FUSED, P, I, FUSED>;                                               \
    PICNIX_CUDA(a, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                        (int)SMEM_BYTES));                                         \
    kern<<<blocks, THREADS, SMEM_BYTES, a->stream>>>(g, a->d, rc, c0, cn, delt);                   \
  } else {                                                                                         \
    auto kern = row_push2d_kernel<FUSED, P, I, false>;                                               \
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
  return check_cuda(a, cudaGetLastError(), "row_push2d_kernel");
}

} // namespace

// 2-D (z ignorable), 2nd-order shapes, rows that split into RX-cell segments and WARPS-row groups
bool row_push2d_geometry(const picnix_arena* a)
{
  const Geom& g = a->g;
  return g.dimension == 2 && g.has_dim[0] == 0 && g.has_dim[1] && g.has_dim[2] && g.order == 2 &&
         (g.dims[2] % rowtile::RX) == 0 && (g.dims[1] % rowtile::WARPS) == 0 && g.Ns <= rowtile::MAXNS;
}

int launch_deposit_rows_2d(picnix_arena* a, int c0, int cn, double delt)
{
  return launch_row_kernel<false>(a, c0, cn, delt);
}

int launch_row_fused_2d(picnix_arena* a, int c0, int cn, double delt)
{
  return launch_row_kernel<true>(a, c0, cn, delt);
}

} // namespace picnix
