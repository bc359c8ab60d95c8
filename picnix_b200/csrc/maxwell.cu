// -*- C++ -*-
// K3: Yee FDTD field update with the Friedman time filter, batched over all chunks of the arena.
//
// Reference: pic_engine::BaseMaxwell (pic/engine/maxwell.hpp)
//   push_bfd_{1,2,3}d  :148-201, 292-348, 442-496   B += -+ cfl * curl(ff0),  ff0 = A E + B ff1 + C ff2
//   push_efd_{1,2,3}d  : 85-146, 231-290, 383-440   ff2 <- ff1 + theta ff2; ff1 <- E; E += +-cfl curl B - dt J
//   init_friedman      : 44-61
//   get_diverror_{1,2,3}d : 63-83, 203-229, 350-381
//   PicChunk::get_energy (field part), pic/pic_chunk.cpp:407-426
//
// This translation unit is compiled with -fmad=false and evaluates every expression in the
// reference's order, so the results are bit-identical to a non-contracting CPU evaluation.
//
// HBM-bound streaming kernels: one thread per padded cell, consecutive threads on consecutive x,
// a warp reads 32 x 48 B contiguous from uf.  Algorithmic traffic per cell: push_bfd reads E (24 B)
// + ff1,ff2 (48 B) and writes ff0 (24 B) + B (24 B RMW = 48 B); push_efd reads B (24) + J (24) +
// ff1,ff2 (48) and writes ff1,ff2 (48) + E (48 RMW).
#include "arena.hpp"

namespace picnix
{

namespace
{

constexpr int MAXWELL_THREADS = 256; // cells per block (and threads)

struct CellIndex {
  int     chunk, iz, iy, ix;
  int64_t cell; // flat padded cell index within the arena
  bool    valid;
};

// Decode the thread into (chunk, iz, iy, ix).  For ignorable dimensions only the single interior
// plane is visited, as the reference's 1-D/2-D loops do (`int iz = lbz;`).
template <int Dim>
__device__ __forceinline__ CellIndex decode_cell(const Geom& g, int c0, int cn)
{
  CellIndex idx;
  const int nx = g.M[2];
  const int ny = Dim >= 2 ? g.M[1] : 1;
  const int nz = Dim >= 3 ? g.M[0] : 1;
  int64_t   t  = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t   per_chunk = (int64_t)nx * ny * nz;
  idx.valid = t < per_chunk * cn;
  if (!idx.valid)
    return idx;
  int     lc = (int)(t / per_chunk);
  int64_t r  = t - (int64_t)lc * per_chunk;
  int     jz = (int)(r / ((int64_t)nx * ny));
  int     r2 = (int)(r - (int64_t)jz * nx * ny);
  int     jy = r2 / nx;
  idx.ix     = r2 - jy * nx;
  idx.iy     = Dim >= 2 ? jy : g.Lb[1];
  idx.iz     = Dim >= 3 ? jz : g.Lb[0];
  idx.chunk  = c0 + lc;
  idx.cell   = (((int64_t)idx.chunk * g.M[0] + idx.iz) * g.M[1] + idx.iy) * g.M[2] + idx.ix;
  return idx;
}

// Shared-memory staging.  A thread owns one padded cell whose uf (48 B) and ff (72 B) records are 48 / 72 B
// apart from its neighbour's: loading them thread by thread touches 12-18 cache lines per warp instruction
// and the kernels end up bound by the L1 tag rate, not by HBM.  The cells of a block are contiguous in
// memory (always in 3-D; inside one chunk's plane / row in 2-D / 1-D), so the block copies its whole range
// with coalesced loads, works in shared memory, and writes the range back with coalesced stores.  Values
// the kernel does not change are written back unchanged (bit-identical, nobody else writes them).
// Neighbours inside the block's range (x-1 / x+1 always but for one thread, y-1 / y+1 mostly) are read
// from the stage, the others from global memory.  Blocks that straddle a gap fall back to direct access.
template <int Dim>
__device__ __forceinline__ int64_t flat_cell(const Geom& g, int c0, int64_t t)
{
  const int     nx = g.M[2];
  const int     ny = Dim >= 2 ? g.M[1] : 1;
  const int     nz = Dim >= 3 ? g.M[0] : 1;
  const int64_t per_chunk = (int64_t)nx * ny * nz;
  const int     lc = (int)(t / per_chunk);
  const int64_t r  = t - (int64_t)lc * per_chunk;
  const int     jz = (int)(r / ((int64_t)nx * ny));
  const int     r2 = (int)(r - (int64_t)jz * nx * ny);
  const int     jy = r2 / nx;
  const int     ix = r2 - jy * nx;
  const int     iy = Dim >= 2 ? jy : g.Lb[1];
  const int     iz = Dim >= 3 ? jz : g.Lb[0];
  return (((int64_t)(c0 + lc) * g.M[0] + iz) * g.M[1] + iy) * g.M[2] + ix;
}

// first cell of the block's contiguous range, or -1 when the block must use direct access (block-uniform)
template <int Dim>
__device__ __forceinline__ int64_t block_range(const Geom& g, int c0, int cn)
{
  const int     nx = g.M[2];
  const int     ny = Dim >= 2 ? g.M[1] : 1;
  const int     nz = Dim >= 3 ? g.M[0] : 1;
  const int64_t total = (int64_t)nx * ny * nz * cn;
  const int64_t t0    = (int64_t)blockIdx.x * MAXWELL_THREADS;
  if (t0 + MAXWELL_THREADS > total)
    return -1;
  const int64_t first = flat_cell<Dim>(g, c0, t0);
  const int64_t last  = flat_cell<Dim>(g, c0, t0 + MAXWELL_THREADS - 1);
  return last - first == MAXWELL_THREADS - 1 ? first : -1;
}

template <int W>
__device__ __forceinline__ void stage_in(double* __restrict__ s, const double* __restrict__ gsrc, int64_t first)
{
#pragma unroll
  for (int i = 0; i < W; i++)
    s[i * MAXWELL_THREADS + threadIdx.x] = gsrc[first * W + i * MAXWELL_THREADS + threadIdx.x];
}

template <int W>
__device__ __forceinline__ void stage_out(double* __restrict__ gdst, const double* __restrict__ s, int64_t first)
{
#pragma unroll
  for (int i = 0; i < W; i++)
    gdst[first * W + i * MAXWELL_THREADS + threadIdx.x] = s[i * MAXWELL_THREADS + threadIdx.x];
}

// record of `cell`: from the stage when it lies in the block's range
template <int W>
__device__ __forceinline__ const double* record(const double* s, const double* gsrc, int64_t first, int64_t cell)
{
  const int64_t li = cell - first;
  return (first >= 0 && li >= 0 && li < MAXWELL_THREADS) ? s + li * W : gsrc + cell * W;
}

// ff(.., 0, k) = A * uf(.., k) + B * ff(.., 1, k) + C * ff(.., 2, k)
__device__ __forceinline__ double filtered(const double* ufr, const double* ffr, int k, double A, double B, double C)
{
  return A * ufr[k] + B * ffr[3 + k] + C * ffr[6 + k];
}

template <int Dim>
__global__ void __launch_bounds__(MAXWELL_THREADS)
push_bfd_kernel(Geom g, DevPtrs d, int c0, int cn, double delt)
{
  __shared__ double s_uf[MAXWELL_THREADS * 6];
  __shared__ double s_ff[MAXWELL_THREADS * 9];

  const int64_t first = block_range<Dim>(g, c0, cn);
  if (first >= 0) {
    stage_in<6>(s_uf, d.uf, first);
    stage_in<9>(s_ff, d.ff, first);
    __syncthreads();
  }
  CellIndex idx = decode_cell<Dim>(g, c0, cn);
  if (!idx.valid)
    return; // never in a staged block (its range is complete)

  const double theta = g.theta;
  const double A     = 1 + 0.5 * theta;
  const double B     = -theta * (1 - 0.5 * theta);
  const double C     = 0.5 * theta * (1 - theta) * (1 - theta);
  const double cflx  = g.cc * delt / g.del[2];
  const double cfly  = g.cc * delt / g.del[1];
  const double cflz  = g.cc * delt / g.del[0];

  const int64_t sx = 1, sy = g.M[2], sz = (int64_t)g.M[1] * g.M[2];
  const int64_t c  = idx.cell;
  // own records (read and written), neighbour records (read only; E, ff1, ff2 are inputs of this kernel)
  double*       ufc = first >= 0 ? s_uf + (c - first) * 6 : d.uf + c * 6;
  double*       ffc = first >= 0 ? s_ff + (c - first) * 9 : d.ff + c * 9;
  const double* ufx = record<6>(s_uf, d.uf, first, c - sx);
  const double* ffx = record<9>(s_ff, d.ff, first, c - sx);
  const double* ufy = record<6>(s_uf, d.uf, first, c - sy);
  const double* ffy = record<9>(s_ff, d.ff, first, c - sy);
  const double* ufz = d.uf + (c - sz) * 6; // a plane away: never in the block's range
  const double* ffz = d.ff + (c - sz) * 9;

  // filtered E at this cell (stored) and at the -1 neighbours (recomputed, same expression)
  double f0x = filtered(ufc, ffc, 0, A, B, C);
  double f0y = filtered(ufc, ffc, 1, A, B, C);
  double f0z = filtered(ufc, ffc, 2, A, B, C);

  // lower bounds of the B loops are lb - Nb + 1 == 1 in the staggered directions
  const bool okx = idx.ix >= 1;
  const bool oky = Dim >= 2 ? idx.iy >= 1 : true;
  const bool okz = Dim >= 3 ? idx.iz >= 1 : true;

  double bx = ufc[3], by = ufc[4], bz = ufc[5];
  if (Dim == 1) {
    if (okx) {
      double fzx = filtered(ufx, ffx, 2, A, B, C);
      double fyx = filtered(ufx, ffx, 1, A, B, C);
      by += (+cflx) * (f0z - fzx);
      bz += (-cflx) * (f0y - fyx);
    }
  } else if (Dim == 2) {
    if (oky) {
      double fzy = filtered(ufy, ffy, 2, A, B, C);
      bx += (-cfly) * (f0z - fzy);
    }
    if (okx) {
      double fzx = filtered(ufx, ffx, 2, A, B, C);
      by += (+cflx) * (f0z - fzx);
    }
    if (okx && oky) {
      double fyx = filtered(ufx, ffx, 1, A, B, C);
      double fxy = filtered(ufy, ffy, 0, A, B, C);
      bz += (-cflx) * (f0y - fyx) + (+cfly) * (f0x - fxy);
    }
  } else {
    if (okz && oky) {
      double fzy = filtered(ufy, ffy, 2, A, B, C);
      double fyz = filtered(ufz, ffz, 1, A, B, C);
      bx += (-cfly) * (f0z - fzy) + (+cflz) * (f0y - fyz);
    }
    if (okz && okx) {
      double fxz = filtered(ufz, ffz, 0, A, B, C);
      double fzx = filtered(ufx, ffx, 2, A, B, C);
      by += (-cflz) * (f0x - fxz) + (+cflx) * (f0z - fzx);
    }
    if (oky && okx) {
      double fyx = filtered(ufx, ffx, 1, A, B, C);
      double fxy = filtered(ufy, ffy, 0, A, B, C);
      bz += (-cflx) * (f0y - fyx) + (+cfly) * (f0x - fxy);
    }
  }
  // outputs: ff0 and B; in a staged block they go to the stage after every thread has read its inputs
  // (ff0 and B are nobody's input here, so no barrier is needed before these stores)
  ffc[0] = f0x;
  ffc[1] = f0y;
  ffc[2] = f0z;
  ufc[3] = bx;
  ufc[4] = by;
  ufc[5] = bz;
  if (first >= 0) {
    __syncthreads();
    stage_out<6>(d.uf, s_uf, first);
    stage_out<9>(d.ff, s_ff, first);
  }
}

template <int Dim>
__global__ void __launch_bounds__(MAXWELL_THREADS)
push_efd_kernel(Geom g, DevPtrs d, int c0, int cn, double delt)
{
  __shared__ double s_uf[MAXWELL_THREADS * 6];
  __shared__ double s_ff[MAXWELL_THREADS * 9];
  __shared__ double s_uj[MAXWELL_THREADS * 4];

  const int64_t first = block_range<Dim>(g, c0, cn);
  if (first >= 0) {
    stage_in<6>(s_uf, d.uf, first);
    stage_in<9>(s_ff, d.ff, first);
    stage_in<4>(s_uj, d.uj, first);
    __syncthreads();
  }
  CellIndex idx = decode_cell<Dim>(g, c0, cn);
  if (!idx.valid)
    return;

  const double theta = g.theta;
  const double cflx  = g.cc * delt / g.del[2];
  const double cfly  = g.cc * delt / g.del[1];
  const double cflz  = g.cc * delt / g.del[0];

  const int64_t sx = 1, sy = g.M[2], sz = (int64_t)g.M[1] * g.M[2];
  const int64_t c  = idx.cell;
  double*       ufc = first >= 0 ? s_uf + (c - first) * 6 : d.uf + c * 6;
  double*       ffc = first >= 0 ? s_ff + (c - first) * 9 : d.ff + c * 9;
  const double* ujc = first >= 0 ? s_uj + (c - first) * 4 : d.uj + c * 4;
  // the +1 neighbours' B (an input of this kernel: nobody writes it)
  const double* ufx = record<6>(s_uf, d.uf, first, c + sx);
  const double* ufy = record<6>(s_uf, d.uf, first, c + sy);
  const double* ufz = d.uf + (c + sz) * 6;

  // Friedman history shift first (uses E before the update)
  double ex = ufc[0], ey = ufc[1], ez = ufc[2];
  ffc[6] = ffc[3] + theta * ffc[6];
  ffc[3] = ex;
  ffc[7] = ffc[4] + theta * ffc[7];
  ffc[4] = ey;
  ffc[8] = ffc[5] + theta * ffc[8];
  ffc[5] = ez;

  // upper bounds of the E loops are ub + Nb - 1 == M - 2 in the staggered directions
  const bool okx = idx.ix <= g.M[2] - 2;
  const bool oky = Dim >= 2 ? idx.iy <= g.M[1] - 2 : true;
  const bool okz = Dim >= 3 ? idx.iz <= g.M[0] - 2 : true;

  const double bx = ufc[3], by = ufc[4], bz = ufc[5];
  // the reads of the neighbours' B precede the barrier below; E (written here) is nobody's input

  if (Dim == 1) {
    ufc[0] = ex + (-delt * ujc[1]);
    if (okx) {
      ufc[1] = ey + ((-cflx) * (ufx[5] - bz) - delt * ujc[2]);
      ufc[2] = ez + ((+cflx) * (ufx[4] - by) - delt * ujc[3]);
    }
  } else if (Dim == 2) {
    if (oky) {
      ufc[0] = ex + ((+cfly) * (ufy[5] - bz) - delt * ujc[1]);
    }
    if (okx) {
      ufc[1] = ey + ((-cflx) * (ufx[5] - bz) - delt * ujc[2]);
    }
    if (okx && oky) {
      ufc[2] = ez + ((+cflx) * (ufx[4] - by) + (-cfly) * (ufy[3] - bx) - delt * ujc[3]);
    }
  } else {
    if (okz && oky) {
      ufc[0] = ex + ((+cfly) * (ufy[5] - bz) + (-cflz) * (ufz[4] - by) - delt * ujc[1]);
    }
    if (okz && okx) {
      ufc[1] = ey + ((+cflz) * (ufz[3] - bx) + (-cflx) * (ufx[5] - bz) - delt * ujc[2]);
    }
    if (oky && okx) {
      ufc[2] = ez + ((+cflx) * (ufx[4] - by) + (-cfly) * (ufy[3] - bx) - delt * ujc[3]);
    }
  }
  if (first >= 0) {
    __syncthreads();
    stage_out<6>(d.uf, s_uf, first);
    stage_out<9>(d.ff, s_ff, first);
  }
}

// init_friedman visits the WHOLE padded array in every dimensionality (maxwell.hpp:50-60)
__global__ void init_friedman_kernel(Geom g, DevPtrs d, int c0, int cn)
{
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)cn * g.Ng)
    return;
  int64_t c = (int64_t)c0 * g.Ng + t;
  for (int k = 0; k < 3; k++) {
    double e         = d.uf[c * 6 + k];
    d.ff[c * 9 + k]     = e;
    d.ff[c * 9 + 3 + k] = e;
    d.ff[c * 9 + 6 + k] = e;
  }
}

// one block per chunk; deterministic tree reduction
template <int Dim, int Mode>
__global__ void __launch_bounds__(256) reduce_kernel(Geom g, DevPtrs d, double* out)
{
  __shared__ double s0[256];
  __shared__ double s1[256];

  const int     chunk = blockIdx.x;
  const int64_t sx = 1, sy = g.M[2], sz = (int64_t)g.M[1] * g.M[2];
  const int*    nbr = d.nbr + chunk * NBSIZE;

  // Maxwell::get_diverror skips the margin next to a physical boundary (pic_engine.hpp:42-71)
  int lo[3], hi[3];
  for (int i = 0; i < 3; i++) {
    lo[i] = g.Lb[i];
    hi[i] = g.Ub[i];
  }
  if (Mode == 0) {
    if (nbr[9 * 1 + 3 * 1 + 0] == NB_NONE) lo[2] += g.nb;
    if (nbr[9 * 1 + 3 * 1 + 2] == NB_NONE) hi[2] -= g.nb;
    if (nbr[9 * 1 + 3 * 0 + 1] == NB_NONE) lo[1] += g.nb;
    if (nbr[9 * 1 + 3 * 2 + 1] == NB_NONE) hi[1] -= g.nb;
    if (nbr[9 * 0 + 3 * 1 + 1] == NB_NONE) lo[0] += g.nb;
    if (nbr[9 * 2 + 3 * 1 + 1] == NB_NONE) hi[0] -= g.nb;
  }
  if (Dim < 3) { lo[0] = hi[0] = g.Lb[0]; }
  if (Dim < 2) { lo[1] = hi[1] = g.Lb[1]; }

  const int nx = hi[2] - lo[2] + 1, ny = hi[1] - lo[1] + 1, nz = hi[0] - lo[0] + 1;
  const int n  = (nx > 0 && ny > 0 && nz > 0) ? nx * ny * nz : 0;

  const double rdx = 1 / g.del[2], rdy = 1 / g.del[1], rdz = 1 / g.del[0];
  const double* uf = d.uf;
  const double* uj = d.uj;

  double acc0 = 0, acc1 = 0;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    int     jz = t / (nx * ny);
    int     r  = t - jz * nx * ny;
    int     jy = r / nx;
    int     jx = r - jy * nx;
    int64_t c  = (((int64_t)chunk * g.M[0] + lo[0] + jz) * g.M[1] + lo[1] + jy) * g.M[2] + lo[2] + jx;
    if (Mode == 0) {
      double dive = (uf[(c + sx) * 6 + 0] - uf[c * 6 + 0]) * rdx;
      double divb = (uf[c * 6 + 3] - uf[(c - sx) * 6 + 3]) * rdx;
      if (Dim >= 2) {
        dive = dive + (uf[(c + sy) * 6 + 1] - uf[c * 6 + 1]) * rdy;
        divb = divb + (uf[c * 6 + 4] - uf[(c - sy) * 6 + 4]) * rdy;
      }
      if (Dim >= 3) {
        dive = dive + (uf[(c + sz) * 6 + 2] - uf[c * 6 + 2]) * rdz;
        divb = divb + (uf[c * 6 + 5] - uf[(c - sz) * 6 + 5]) * rdz;
      }
      acc0 += dive - uj[c * 4 + 0];
      acc1 += divb;
    } else {
      // energy sums run over the full interior in every dimensionality (pic_chunk.cpp:415-426)
      double ex = uf[c * 6 + 0], ey = uf[c * 6 + 1], ez = uf[c * 6 + 2];
      double bx = uf[c * 6 + 3], by = uf[c * 6 + 4], bz = uf[c * 6 + 5];
      acc0 += 0.5 * (ex * ex + ey * ey + ez * ez);
      acc1 += 0.5 * (bx * bx + by * by + bz * bz);
    }
  }

  s0[threadIdx.x] = acc0;
  s1[threadIdx.x] = acc1;
  __syncthreads();
  for (int w = blockDim.x / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) {
      s0[threadIdx.x] += s0[threadIdx.x + w];
      s1[threadIdx.x] += s1[threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[chunk * 2 + 0] = s0[0];
    out[chunk * 2 + 1] = s1[0];
  }
}

template <int Dim>
int64_t cells_visited(const Geom& g, int cn)
{
  int64_t nx = g.M[2];
  int64_t ny = Dim >= 2 ? g.M[1] : 1;
  int64_t nz = Dim >= 3 ? g.M[0] : 1;
  return nx * ny * nz * cn;
}

} // namespace

int launch_init_friedman(picnix_arena* a, int c0, int cn)
{
  resolve_range(a, c0, cn);
  int64_t n = (int64_t)cn * a->g.Ng;
  if (n == 0)
    return PICNIX_OK;
  int blocks = (int)((n + 255) / 256);
  init_friedman_kernel<<<blocks, 256, 0, a->stream>>>(a->g, a->d, c0, cn);
  a->kernel_launches++;
  return check_cuda(a, cudaGetLastError(), "init_friedman");
}

int launch_push_bfd(picnix_arena* a, int c0, int cn, double delt)
{
  resolve_range(a, c0, cn);
  if (cn == 0)
    return PICNIX_OK;
  const Geom& g = a->g;
  int64_t     n;
  switch (g.dimension) {
  case 1:
    n = cells_visited<1>(g, cn);
    push_bfd_kernel<1><<<(int)((n + MAXWELL_THREADS - 1) / MAXWELL_THREADS), MAXWELL_THREADS, 0,
                         a->stream>>>(g, a->d, c0, cn, delt);
    break;
  case 2:
    n = cells_visited<2>(g, cn);
    push_bfd_kernel<2><<<(int)((n + MAXWELL_THREADS - 1) / MAXWELL_THREADS), MAXWELL_THREADS, 0,
                         a->stream>>>(g, a->d, c0, cn, delt);
    break;
  default:
    n = cells_visited<3>(g, cn);
    push_bfd_kernel<3><<<(int)((n + MAXWELL_THREADS - 1) / MAXWELL_THREADS), MAXWELL_THREADS, 0,
                         a->stream>>>(g, a->d, c0, cn, delt);
    break;
  }
  a->kernel_launches++;
  return check_cuda(a, cudaGetLastError(), "push_bfd");
}

int launch_push_efd(picnix_arena* a, int c0, int cn, double delt)
{
  resolve_range(a, c0, cn);
  if (cn == 0)
    return PICNIX_OK;
  const Geom& g = a->g;
  int64_t     n;
  switch (g.dimension) {
  case 1:
    n = cells_visited<1>(g, cn);
    push_efd_kernel<1><<<(int)((n + MAXWELL_THREADS - 1) / MAXWELL_THREADS), MAXWELL_THREADS, 0,
                         a->stream>>>(g, a->d, c0, cn, delt);
    break;
  case 2:
    n = cells_visited<2>(g, cn);
    push_efd_kernel<2><<<(int)((n + MAXWELL_THREADS - 1) / MAXWELL_THREADS), MAXWELL_THREADS, 0,
                         a->stream>>>(g, a->d, c0, cn, delt);
    break;
  default:
    n = cells_visited<3>(g, cn);
    push_efd_kernel<3><<<(int)((n + MAXWELL_THREADS - 1) / MAXWELL_THREADS), MAXWELL_THREADS, 0,
                         a->stream>>>(g, a->d, c0, cn, delt);
    break;
  }
  a->kernel_launches++;
  return check_cuda(a, cudaGetLastError(), "push_efd");
}

template <int Mode>
static int launch_reduce(picnix_arena* a, double* out0, double* out1)
{
  const Geom& g = a->g;
  switch (g.dimension) {
  case 1:
    reduce_kernel<1, Mode><<<g.nchunk, 256, 0, a->stream>>>(g, a->d, a->d_reduce);
    break;
  case 2:
    reduce_kernel<2, Mode><<<g.nchunk, 256, 0, a->stream>>>(g, a->d, a->d_reduce);
    break;
  default:
    reduce_kernel<3, Mode><<<g.nchunk, 256, 0, a->stream>>>(g, a->d, a->d_reduce);
    break;
  }
  a->kernel_launches++;
  PICNIX_CUDA(a, cudaGetLastError());
  std::vector<double> host((size_t)g.nchunk * 2);
  PICNIX_CUDA(a, cudaMemcpyAsync(host.data(), a->d_reduce, host.size() * sizeof(double),
                                 cudaMemcpyDeviceToHost, a->stream));
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  for (int i = 0; i < g.nchunk; i++) {
    out0[i] = host[2 * i + 0];
    out1[i] = host[2 * i + 1];
  }
  return PICNIX_OK;
}

int launch_diverror(picnix_arena* a, double* efd, double* bfd)
{
  return launch_reduce<0>(a, efd, bfd);
}

int launch_field_energy(picnix_arena* a, double* efd, double* bfd)
{
  return launch_reduce<1>(a, efd, bfd);
}

} // namespace picnix
