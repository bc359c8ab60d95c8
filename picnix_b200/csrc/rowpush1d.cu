// -*- C++ -*-
// Tiled kernel for 1-D runs (x; y, z ignorable), order 2: the one-dimensional variant of rowpush.cu, see
// rowtile1d.cuh for the scheme.  One warp per (chunk, x-segment of RX cells); a block is WARPS consecutive
// segments of the chunk range (they may belong to different chunks: a two-stream chunk is one segment), so
// the field tile and the segment offsets are per warp.  FUSED / PERM as in rowpush.cu.
#include "rowtile1d.cuh"

namespace picnix
{

namespace
{

using namespace rowtile1d;

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
    esirkepov_deposit<1, 2>(g, lim, r[6], delt, r[0], r[1], r[2], r[3], r[4], r[5], bz, by, bx, add);
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
__global__ void __launch_bounds__(THREADS, 4)
row_push1d_kernel(Geom g, DevPtrs d, RowConst rc, int c0, int cn, double delt)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BlockSmem* bs  = reinterpret_cast<BlockSmem*>(smem_raw);
  WarpSmem*  wsm = reinterpret_cast<WarpSmem*>(smem_raw + sizeof(BlockSmem));

  const int      lane = threadIdx.x & 31;
  const int      warp = threadIdx.x >> 5;
  const int      half = lane >> 4;
  const unsigned FULL = 0xffffffffu;
  WarpSmem*      ws   = wsm + warp;
  const LaneMap  lm   = lane_map(lane);
  const int      Ns   = g.Ns;

  // warp -> (chunk, x segment); the warps beyond the last segment walk an empty stream
  const int  nsegx  = g.dims[2] / RX;
  const int  sidx   = blockIdx.x * WARPS + warp;
  const bool active = sidx < nsegx * cn;
  const int  sclamp = active ? sidx : nsegx * cn - 1;
  const int  lc     = sclamp / nsegx;
  const int  jx0    = (sclamp - lc * nsegx) * RX;
  const int  chunk  = c0 + lc;

  const double* lim = d.clim + chunk * 6;
  double*       uj  = d.uj + (int64_t)chunk * g.Ng * 4;
  const int     My = g.M[1], Mx = g.M[2];

  // ---- the field tile of the segment starts travelling (layout [x][6], 16-byte asynchronous copies) ----
  if (FUSED) {
    const double* uf = d.uf + (int64_t)chunk * g.Ng * 6;
    const int     gz = g.Lb[0], gy = g.Lb[1], gx = jx0 + g.Lb[2] - 1;
    for (int e = lane; e < FTILE / 2; e += 32)
      cp_async_16(reinterpret_cast<double2*>(ws->ftile) + e,
                  reinterpret_cast<const double2*>(uf + ((int64_t)(gz * My + gy) * Mx + gx) * 6) + e);
  }
  if (threadIdx.x < Ns) {
    const int    is = threadIdx.x;
    const double q  = d.qm[2 * is];
    bs->q[is]       = q;
    bs->qmdt[is]    = 0.5 * q / d.qm[2 * is + 1] * delt;
  }
  if (lane < Ns)
    ws->off[lane] = d.seg_off[chunk * Ns + lane];

  // ---- the stream of this warp's segment: lane (cell c, species is) = c * Ns + is builds its entry;
  // cells are padded to a multiple of ALIGN slots ----
  const int key0 = jx0;
  const int nent = RX * Ns;
  {
    const int c  = lane / Ns;
    const int is = lane - c * Ns;
    int       b = 0, n = 0;
    if (lane < nent && active) {
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
  for (int i = lane; i < REC; i += 32)
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
      const int64_t off = ws->off[cur.sc & 0xff];
      const int64_t i   = PERM ? off + d.perm[off + cur.idx] : off + cur.idx;
#pragma unroll
      for (int k = 0; k < (PERM ? 7 : 6); k++)
        cp_async_f64(&ws->pfb[k][lane], d.xu + k * d.pcap + i);
    }
    if (PERM && nxt.idx >= 0)
      cp_async_i32(ws->pbuf + lane, d.perm + ws->off[nxt.sc & 0xff] + nxt.idx);
    cp_async_commit_wait();
  }

  const double rdx = rc.rd[2];
  const double dx  = rc.del[2];
  // chunk limits and the first cell-centre point, pic/engine/velocity.hpp:304-315
  if (lane == 0) {
    ws->rowc[0] = lim[4];
    ws->rowc[1] = lim[5];
    ws->rowc[2] = lim[4] + 0.5 * dx;
  }
  __syncwarp();
  const double xmin = ws->rowc[0], xmax = ws->rowc[1], xigrid = ws->rowc[2];

  double acc     = 0;
  int    curinfo = -1; // info word of the cell the accumulator belongs to (-1: none)

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
    __syncwarp(); // every lane has its values before the buffer is refilled
    // the next batch starts travelling now and has phases 1 and 2 of this one to arrive; the permutation
    // entries are requested two batches ahead
    const Slot nn = stream_slot(ws, base + 64 + lane, nent, kent);
    if (FUSED) {
      if (nxt.idx >= 0) {
        const int64_t off = ws->off[nxt.sc & 0xff];
        const int64_t i   = PERM ? off + ws->pbuf[lane] : off + nxt.idx;
#pragma unroll
        for (int k = 0; k < (PERM ? 7 : 6); k++)
          cp_async_f64(&ws->pfb[k][lane], d.xu + k * d.pcap + i);
      }
      if (PERM && nn.idx >= 0)
        cp_async_i32(ws->pbuf + lane, d.perm + ws->off[nn.sc & 0xff] + nn.idx);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }

    // ---------------- phase 1: one particle per lane ----------------
    int inf = 0;
    if (cur.idx >= 0) {
      const int     is = cur.sc & 0xff;
      const int     jx = cur.sc >> 8;  // old cell relative to the segment: given by the sort
      const int     cx = jx0 + jx;
      const int64_t i  = ws->off[is] + cur.idx;
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

        // weights on the centre grid (MC or WT) and on the edge grid (MC); the particle is in cell cx by
        // construction of the sort
        double       s0x[3], wix[3], h[3], whx[4];
        const double dix = (x0 - (xigrid + cxf * dx)) * rdx;
        shape2(dix, s0x);
        if (Interp == PICNIX_INTERP_MC) {
#pragma unroll
          for (int k = 0; k < 3; k++)
            wix[k] = s0x[k];
        } else {
          shape_wt<2>(x0, xigrid + cxf * dx, rdx, rc.cfl[2], 1 / rc.cfl[2], wix);
        }
        // nearest cell edge: the one to the right when the particle sits right of the centre; its
        // three weights go into the cell-anchored 4-slot array
        const bool hx = dix >= 0.0;
        shape2((x0 - (xmin + (cxf + (hx ? 1.0 : 0.0)) * dx)) * rdx, h);
        shift4(h, hx, whx);

        // first stencil point of the cell in the tile; Yee staggering, pic/engine/velocity.hpp:376-381
        const double* F    = ws->ftile + jx * 6;
        const double  qmdt = bs->qmdt[is];
        double ex = interp_cell<4>(F + 0, whx) * qmdt;
        double ey = interp_cell<3>(F + 1, wix) * qmdt;
        double ez = interp_cell<3>(F + 2, wix) * qmdt;
        double bx = interp_cell<3>(F + 3, wix) * qmdt;
        double by = interp_cell<4>(F + 4, whx) * qmdt;
        double bz = interp_cell<4>(F + 5, whx) * qmdt;

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
      if (FUSED) {
        const int seg = chunk * Ns + is;
        const int key = (x1 < xmin || x1 >= xmax) ? g.Ng : ix1;
        d.gindex[i]   = key;
        atomicAdd(d.pcount + (int64_t)seg * (g.Ng + 1) + key, 1);
        if (key == g.Ng)
          note_leaver(d, seg, i);
      }

      double s0x[3], s1x[3];
      shape2((x0 - (xigrid + cxf * dx)) * rdx, s0x);
      shape2((x1 - (xigrid + (double)ix1 * dx)) * rdx, s1x);
      const int shx = ix1 - cx;
      if (abs(shx) <= 1) {
        const AxisFactors fx = window_factors(s0x, s1x, shx);
        // the ignorable directions contribute through the velocity (pic/engine/current.hpp:220-221)
        const double qvy = q * ((y1 - y0) / delt);
        const double qvz = q * ((z1 - z0) / delt);
        stage_particle(ws->stg + lane * REC, fx, q, qvy, qvz, rc.ddt[2]);
        inf = make_info(jx, fx.w, 1, 1);
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
          acc = 0;
        }
        curinfo = ginfo;
      }
      const int cnt = last - L + 1;
      if (__popc(gm) == cnt) {
        // no foreign slot inside the range (the usual case): plain pointer walk
        const double* rec = ws->stg + (L + half) * REC;
#pragma unroll 2
        for (int k = 0; k < (cnt >> 1); k++) {
          acc += rec[lm.v];
          rec += 2 * REC;
        }
        if (cnt & 1)
          acc += (half == 0 ? rec : ws->zero)[lm.v];
      } else {
#pragma unroll 1
        for (int j = L + half; j <= last + half; j += 2) {
          const double* rec = ((gm >> (j & 31)) & 1u) && j <= last ? ws->stg + j * REC : ws->zero;
          acc += rec[lm.v];
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

  // warp tile -> global current: one fp64 reduction per non-zero tile value; the tile is one contiguous
  // run of uj (the four components of a point stay together in both)
  {
    const int     gz0 = g.Lb[0], gy0 = g.Lb[1], gx0 = jx0 + g.Lb[2] - 2;
    double*       dst = uj + ((int64_t)(gz0 * My + gy0) * Mx + gx0) * 4;
    for (int e = lane; e < TILE; e += 32) {
      const double v = ws->tile[e];
      if (v != 0.0)
        atomicAdd(dst + e, v);
    }
  }
}

template <bool FUSED>
int launch_row_kernel(picnix_arena* a, int c0, int cn, double delt)
{
  const Geom& g      = a->g;
  const int   blocks = ((g.dims[2] / RX) * cn + WARPS - 1) / WARPS;
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
    auto kern = row_push1d_kernel<FUSED, P, I, FUSED>;                                               \
    PICNIX_CUDA(a, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                        (int)SMEM_BYTES));                                         \
    kern<<<blocks, THREADS, SMEM_BYTES, a->stream>>>(g, a->d, rc, c0, cn, delt);                   \
  } else {                                                                                         \
    auto kern = row_push1d_kernel<FUSED, P, I, false>;                                               \
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
  return check_cuda(a, cudaGetLastError(), "row_push1d_kernel");
}

} // namespace

// 1-D (y, z ignorable), 2nd-order shapes, rows that split into RX-cell segments
bool row_push1d_geometry(const picnix_arena* a)
{
  const Geom& g = a->g;
  return g.dimension == 1 && g.has_dim[0] == 0 && g.has_dim[1] == 0 && g.has_dim[2] && g.order == 2 &&
         (g.dims[2] % rowtile::RX) == 0 && g.Ns <= rowtile::MAXNS;
}

int launch_deposit_rows_1d(picnix_arena* a, int c0, int cn, double delt)
{
  return launch_row_kernel<false>(a, c0, cn, delt);
}

int launch_row_fused_1d(picnix_arena* a, int c0, int cn, double delt)
{
  return launch_row_kernel<true>(a, c0, cn, delt);
}

} // namespace picnix
