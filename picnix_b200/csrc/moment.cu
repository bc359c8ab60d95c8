// -*- C++ -*-
// Velocity moments and particle energy: PicChunk::deposit_moment (pic/pic_chunk.cpp:525-534,
// pic/engine/moment.hpp), the particle part of PicChunk::get_energy (pic/pic_chunk.cpp:428-438) and
// the BoundaryMom halo for neighbours inside the arena (nix/xtensor_halo3d.hpp:134-185).
//
//   um[chunk][Mz][My][Mx][Ns][14]   component order pic/engine/moment.hpp:19-32
//       0 t   1 x   2 y   3 z   4 tt   5 xx   6 yy   7 zz   8 tx   9 ty   10 tz   11 xy   12 yz   13 zx
//   per particle and stencil point (momentum-conserving shape, (Order+1)^Dim points):
//       ww = m * wx * wy * wz,  gamma = sqrt(1 + u^2/c^2)
//       t += ww, (x,y,z) += ww u/gamma, (tx,ty,tz) += ww u, tt += ww gamma c, pairs += ww u_i u_j / gamma
//
// Diagnostics cadence, not the per-step hot path.  Two kernels:
//   * moment_row_kernel   -- cell-ordered particles, even orders with (Order+1)^Dim <= 32 (the
//     BASELINE configurations): a warp per run of cells, per-particle factors staged once in shared
//     memory, lane = stencil point for the accumulation, a private shared tile per run (see below).
//   * moment_generic_kernel -- anything else: thread per particle, fp64 reductions to global.
// Results differ from the reference by summation order only (tests: 1e-12 of max |um|).
#include "particle_common.cuh"

#include <algorithm>

namespace picnix
{

namespace
{

constexpr int MTHREADS = 128;
constexpr int NMOM     = 14;

// rg = 1/gamma is formed once; the reference divides by gamma in every term, which differs by an ulp
// (far inside the 1e-12 summation-order tolerance) and costs ten fp64 divisions per particle and point
__device__ __forceinline__ void moment_terms(double ww, double ux, double uy, double uz, double gm,
                                             double rg, double cc, double* t)
{
  const double wg = ww * rg;
  t[0]  = ww;
  t[1]  = wg * ux;
  t[2]  = wg * uy;
  t[3]  = wg * uz;
  t[4]  = ww * gm * cc;
  t[5]  = wg * ux * ux;
  t[6]  = wg * uy * uy;
  t[7]  = wg * uz * uz;
  t[8]  = ww * ux;
  t[9]  = ww * uy;
  t[10] = ww * uz;
  t[11] = wg * ux * uy;
  t[12] = wg * uy * uz;
  t[13] = wg * uz * ux;
}

// weights and first stencil index of one axis (BaseMoment::local{1,2,3}d, moment.hpp:219-395)
template <int Order>
__device__ __forceinline__ int moment_axis(double x, double xmin, double dx, int lb, double* w)
{
  const double rdx = 1 / dx;
  const int    ix  = digitize(x, xmin + 0.5 * dx * (Order % 2), rdx);
  shape_mc<Order>(x, xmin + 0.5 * dx + (double)ix * dx, rdx, w);
  return ix + lb - Order / 2;
}

template <int Dim, int Order>
__global__ void __launch_bounds__(MTHREADS)
moment_generic_kernel(Geom g, DevPtrs d, int blocks_per_seg)
{
  const int seg = blockIdx.x / blocks_per_seg;
  const int ip  = (blockIdx.x - seg * blocks_per_seg) * blockDim.x + threadIdx.x;
  if (ip >= d.np[seg])
    return;
  const int     chunk = seg / g.Ns, is = seg - chunk * g.Ns;
  const int64_t i   = d.seg_off[seg] + ip;
  const double* lim = d.clim + chunk * 6;
  const double  ms  = d.qm[2 * is + 1];
  constexpr int N   = Order + 1;

  const double ux = d.xu[3 * d.pcap + i], uy = d.xu[4 * d.pcap + i], uz = d.xu[5 * d.pcap + i];
  const double rc = 1 / g.cc;
  const double gm = sqrt(1 + (ux * ux + uy * uy + uz * uz) * rc * rc);

  double wx[N], wy[N], wz[N];
  int    ix0, iy0 = g.Lb[1], iz0 = g.Lb[0];
  ix0 = moment_axis<Order>(d.xu[0 * d.pcap + i], lim[4], g.del[2], g.Lb[2], wx);
  if (Dim >= 2)
    iy0 = moment_axis<Order>(d.xu[1 * d.pcap + i], lim[2], g.del[1], g.Lb[1], wy);
  if (Dim >= 3)
    iz0 = moment_axis<Order>(d.xu[2 * d.pcap + i], lim[0], g.del[0], g.Lb[0], wz);

  double* um = d.um + (int64_t)chunk * g.Ng * g.Ns * NMOM;
  for (int jz = 0; jz < (Dim >= 3 ? N : 1); jz++)
    for (int jy = 0; jy < (Dim >= 2 ? N : 1); jy++)
      for (int jx = 0; jx < N; jx++) {
        double ww = ms * wx[jx];
        if (Dim >= 2)
          ww = ww * wy[jy];
        if (Dim >= 3)
          ww = ww * wz[jz];
        double t[NMOM];
        moment_terms(ww, ux, uy, uz, gm, 1 / gm, g.cc, t);
        double* m = um + ((((int64_t)(iz0 + jz) * g.M[1] + iy0 + jy) * g.M[2] + ix0 + jx) * g.Ns + is) * NMOM;
#pragma unroll
        for (int k = 0; k < NMOM; k++)
          atomicAdd(m + k, t[k]);
      }
}

// ------------------------------------------------------------------------------------------------
// Row-tile kernel (cell-ordered particles, even orders with (Order+1)^Dim <= 32: the BASELINE
// configurations).  One warp owns a run of MRUN cells of one row and one species:
//   phase A  lane = particle (coalesced loads, next batch prefetched into registers): gamma, the N weights
//            per axis and the 14 per-particle factors go to a shared record -- formed once per particle;
//   phase B  lane = stencil point: ww = wx[px] wy[py] wz[pz] and 14 FMAs per particle into registers
//            (the record is read with warp-uniform loads);
//   per cell the 14 sums of every point are added to the warp's shared tile (plain adds, the tile is
//   private), and the tile goes to `um` once per run with one reduction per tile element: about
//   (MRUN + N - 1) / MRUN * N^(Dim-1) * 14 global reductions per cell instead of N^Dim * 14.
// ------------------------------------------------------------------------------------------------
constexpr int MRUN   = 8;
constexpr int MWARPS = MTHREADS / 32;
constexpr int MPS    = NMOM + 1; // tile stride of a point: odd number of doubles (no bank conflicts)

template <int Dim, int Order>
struct MomentTile {
  static constexpr int N    = Order + 1;
  static constexpr int NY   = Dim >= 2 ? N : 1;
  static constexpr int NZ   = Dim >= 3 ? N : 1;
  static constexpr int NPTS = N * NY * NZ;
  static constexpr int XT   = MRUN + N - 1;
  static constexpr int TILE = XT * NY * NZ * MPS;
  static constexpr int REC  = 3 * N + NMOM + ((3 * N + NMOM) % 2 == 0 ? 1 : 0); // odd record stride
  static constexpr size_t BYTES = sizeof(double) * MWARPS * (TILE + 32 * REC);
};

template <int Dim, int Order>
__global__ void __launch_bounds__(MTHREADS)
moment_row_kernel(Geom g, DevPtrs d, const int* __restrict__ perm)
{
  // perm != nullptr: a lazy sort is pending -- sorted slot j of a segment still sits in slot perm[j] of xu
  // (sort.cu); the particles are fetched through it instead of being physically reordered first
  using T = MomentTile<Dim, Order>;
  constexpr int N = T::N, NY = T::NY, NZ = T::NZ, NPTS = T::NPTS, XT = T::XT, TILE = T::TILE, REC = T::REC;
  extern __shared__ double msm[];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double*   tile = msm + warp * (TILE + 32 * REC);
  double*   rec  = tile + TILE;

  // warp -> (segment, z, y, run of MRUN cells in x)
  const int     nrun = (g.dims[2] + MRUN - 1) / MRUN;
  const int     ny   = Dim >= 2 ? g.dims[1] : 1;
  const int     nz   = Dim >= 3 ? g.dims[0] : 1;
  const int64_t w    = (int64_t)blockIdx.x * MWARPS + warp;
  const int64_t per_seg = (int64_t)nz * ny * nrun;
  if (w >= per_seg * g.nchunk * g.Ns)
    return;
  const int seg = (int)(w / per_seg);
  int       r   = (int)(w - (int64_t)seg * per_seg);
  const int jz  = r / (ny * nrun);
  r -= jz * ny * nrun;
  const int jy  = r / nrun;
  const int x0  = (r - jy * nrun) * MRUN;
  const int len = min(MRUN, g.dims[2] - x0);

  const int     chunk = seg / g.Ns, is = seg - chunk * g.Ns;
  const int64_t off = d.seg_off[seg];
  const double* lim = d.clim + chunk * 6;
  const double  ms  = d.qm[2 * is + 1];
  const double  rc  = 1 / g.cc;

  // particles of the run: [pix[key0], pix[key0 + len]) with the cell boundaries in between
  const int* pix  = d.pindex + (int64_t)seg * (g.Ng + 1) + (Dim >= 3 ? jz * g.fsz : 0) + (Dim >= 2 ? jy * g.fsy : 0) + x0;
  const int  pcut = lane <= len ? pix[lane] : 0;
  const int  pbeg = __shfl_sync(0xffffffffu, pcut, 0), pend = __shfl_sync(0xffffffffu, pcut, len);
  if (pend <= pbeg)
    return;

  for (int e = lane; e < TILE; e += 32)
    tile[e] = 0;

  const int px = lane % N, py = Dim >= 2 ? (lane / N) % N : 0, pz = Dim >= 3 ? lane / (N * N) : 0;
  double    acc[NMOM];
#pragma unroll
  for (int k = 0; k < NMOM; k++)
    acc[k] = 0;

  // the stream of batches: at most 32 particles, never across a cell boundary
  int    c = 0, base = pbeg; // current cell of the run and first particle of the current batch
  double nx[6];
  auto   fetch = [&](int cell, int b0) {
    const int pe = __shfl_sync(0xffffffffu, pcut, cell + 1);
    if (b0 + lane < pe) {
      const int64_t i = perm != nullptr ? off + perm[off + b0 + lane] : off + b0 + lane;
#pragma unroll
      for (int k = 0; k < 6; k++)
        nx[k] = d.xu[k * d.pcap + i];
    }
  };
  while (c < len && __shfl_sync(0xffffffffu, pcut, c + 1) <= base)
    c++; // leading empty cells
  fetch(c, base);

  while (base < pend) {
    const int pe = __shfl_sync(0xffffffffu, pcut, c + 1);
    const int n  = min(32, pe - base);
    // ---- phase A: one particle per lane ----
    if (lane < n) {
      const double ux = nx[3], uy = nx[4], uz = nx[5];
      const double gm = sqrt(1 + (ux * ux + uy * uy + uz * uz) * rc * rc);
      const double rg = 1 / gm;
      double*      q  = rec + lane * REC;
      double       wgt[N];
      moment_axis<Order>(nx[0], lim[4], g.del[2], g.Lb[2], wgt);
#pragma unroll
      for (int k = 0; k < N; k++)
        q[k] = ms * wgt[k];
      if (Dim >= 2) {
        moment_axis<Order>(nx[1], lim[2], g.del[1], g.Lb[1], wgt);
#pragma unroll
        for (int k = 0; k < N; k++)
          q[N + k] = wgt[k];
      }
      if (Dim >= 3) {
        moment_axis<Order>(nx[2], lim[0], g.del[0], g.Lb[0], wgt);
#pragma unroll
        for (int k = 0; k < N; k++)
          q[2 * N + k] = wgt[k];
      }
      moment_terms(1.0, ux, uy, uz, gm, rg, g.cc, q + 3 * N);
    }
    // next batch: same cell, or the next non-empty one
    int nc = c, nbase = base + n;
    while (nc < len && __shfl_sync(0xffffffffu, pcut, nc + 1) <= nbase)
      nc++;
    if (nbase < pend)
      fetch(nc, nbase);
    __syncwarp();
    // ---- phase B: one stencil point per lane ----
    if (lane < NPTS) {
      for (int p = 0; p < n; p++) {
        const double* q  = rec + p * REC;
        double        ww = q[px];
        if (Dim >= 2)
          ww = ww * q[N + py];
        if (Dim >= 3)
          ww = ww * q[2 * N + pz];
#pragma unroll
        for (int k = 0; k < NMOM; k++)
          acc[k] += ww * q[3 * N + k];
      }
    }
    __syncwarp();
    if (nc != c || nbase >= pend) {
      // the cell is complete: its sums go to the tile
      if (lane < NPTS) {
        double* t = tile + (((c + px) * NY + py) * NZ + pz) * MPS;
#pragma unroll
        for (int k = 0; k < NMOM; k++) {
          t[k] += acc[k];
          acc[k] = 0;
        }
      }
      __syncwarp();
    }
    c    = nc;
    base = nbase;
  }

  // tile -> um; first point of the run's first cell: (jz, jy, x0) + lb - Order/2
  double*   um  = d.um + (int64_t)chunk * g.Ng * g.Ns * NMOM;
  const int ix0 = x0 + g.Lb[2] - Order / 2;
  const int iy0 = Dim >= 2 ? jy + g.Lb[1] - Order / 2 : g.Lb[1];
  const int iz0 = Dim >= 3 ? jz + g.Lb[0] - Order / 2 : g.Lb[0];
  for (int e = lane; e < XT * NY * NZ * NMOM; e += 32) {
    const int    pt = e / NMOM, k = e - pt * NMOM;
    const double v  = tile[pt * MPS + k];
    if (v != 0.0) {
      const int tx = pt / (NY * NZ), ty = (pt / NZ) % NY, tz = pt % NZ;
      atomicAdd(um + ((((int64_t)(iz0 + tz) * g.M[1] + iy0 + ty) * g.M[2] + ix0 + tx) * g.Ns + is) * NMOM + k, v);
    }
  }
}

// BoundaryMom between chunks of the arena: interior-margin cells add the neighbours' ghost cells,
// directions visited in the reference's unpack order (same structure as the current halo)
__global__ void __launch_bounds__(256) moment_halo_local_kernel(Geom g, DevPtrs d)
{
  const int     ncomp = g.Ns * NMOM;
  const int64_t e     = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)g.nchunk * g.Ng * ncomp)
    return;
  const int64_t t     = e / ncomp;
  const int     k     = (int)(e - t * ncomp);
  const int     chunk = (int)(t / g.Ng);
  int           r     = (int)(t - (int64_t)chunk * g.Ng);
  int           idx[3];
  idx[0] = r / (g.M[1] * g.M[2]);
  r -= idx[0] * g.M[1] * g.M[2];
  idx[1] = r / g.M[2];
  idx[2] = r - idx[1] * g.M[2];

  bool reach[3][3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    if (idx[a] < g.Lb[a] || idx[a] > g.Ub[a])
      return;
    reach[a][0] = g.has_dim[a] && idx[a] < g.Lb[a] + g.nb;
    reach[a][1] = true;
    reach[a][2] = g.has_dim[a] && idx[a] > g.Ub[a] - g.nb;
  }
  double acc     = d.um[t * ncomp + k];
  bool   touched = false;
  for (int dz = 0; dz < 3; dz++) {
    if (!reach[0][dz])
      continue;
    for (int dy = 0; dy < 3; dy++) {
      if (!reach[1][dy])
        continue;
      for (int dx = 0; dx < 3; dx++) {
        if (!reach[2][dx] || (dz == 1 && dy == 1 && dx == 1))
          continue;
        const int nb = d.nbr[chunk * NBSIZE + 9 * dz + 3 * dy + dx];
        if (nb < 0)
          continue;
        const int     sz = idx[0] + (dz == 0 ? g.dims[0] : (dz == 2 ? -g.dims[0] : 0));
        const int     sy = idx[1] + (dy == 0 ? g.dims[1] : (dy == 2 ? -g.dims[1] : 0));
        const int     sx = idx[2] + (dx == 0 ? g.dims[2] : (dx == 2 ? -g.dims[2] : 0));
        const int64_t s  = (((int64_t)nb * g.M[0] + sz) * g.M[1] + sy) * g.M[2] + sx;
        acc += d.um[s * ncomp + k];
        touched = true;
      }
    }
  }
  if (touched)
    d.um[t * ncomp + k] = acc;
}

// particle[is] = sum over interior cells of um[..][is][4] * c - um[..][is][0] * c^2, per chunk
__global__ void __launch_bounds__(256) particle_energy_kernel(Geom g, DevPtrs d, double* out)
{
  __shared__ double red[256];
  const int chunk = blockIdx.x / g.Ns, is = blockIdx.x - chunk * g.Ns;
  const int nz = g.Ub[0] - g.Lb[0] + 1, ny = g.Ub[1] - g.Lb[1] + 1, nx = g.Ub[2] - g.Lb[2] + 1;
  double    sum = 0;
  for (int c = threadIdx.x; c < nz * ny * nx; c += blockDim.x) {
    const int     iz = c / (ny * nx) + g.Lb[0], iy = (c / nx) % ny + g.Lb[1], ix = c % nx + g.Lb[2];
    const double* m  = d.um + ((((int64_t)chunk * g.M[0] + iz) * g.M[1] + iy) * g.M[2] + ix) * g.Ns * NMOM +
                      (int64_t)is * NMOM;
    sum += m[4] * g.cc - m[0] * g.cc * g.cc;
  }
  red[threadIdx.x] = sum;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s)
      red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0)
    out[blockIdx.x] = red[0];
}

template <int Dim, int Order>
void launch_moment_kernels(picnix_arena* a)
{
  const Geom&   g    = a->g;
  constexpr int N    = Order + 1;
  constexpr int NPTS = Dim == 3 ? N * N * N : (Dim == 2 ? N * N : N);
  if (a->pindex_valid && Order % 2 == 0 && NPTS <= 32 && !a->force_generic) {
    using T = MomentTile<Dim, Order>;
    auto          kern  = moment_row_kernel<Dim, Order>;
    const int     nrun  = (g.dims[2] + MRUN - 1) / MRUN;
    const int64_t warps = (int64_t)a->nseg * (Dim >= 3 ? g.dims[0] : 1) * (Dim >= 2 ? g.dims[1] : 1) * nrun;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::BYTES);
    kern<<<(unsigned)((warps + MWARPS - 1) / MWARPS), MTHREADS, T::BYTES, a->stream>>>(
        g, a->d, a->perm_pending ? a->d.perm : nullptr);
  } else {
    int maxcap = 0;
    for (int s = 0; s < a->nseg; s++)
      maxcap = std::max(maxcap, a->seg_cap[s]);
    const int bps = (maxcap + MTHREADS - 1) / MTHREADS;
    if (bps > 0)
      moment_generic_kernel<Dim, Order><<<bps * a->nseg, MTHREADS, 0, a->stream>>>(g, a->d, bps);
  }
  a->kernel_launches++;
}

template <int Dim>
void launch_moment_order(picnix_arena* a)
{
  switch (a->g.order) {
  case 1:
    launch_moment_kernels<Dim, 1>(a);
    break;
  case 2:
    launch_moment_kernels<Dim, 2>(a);
    break;
  case 3:
    launch_moment_kernels<Dim, 3>(a);
    break;
  default:
    launch_moment_kernels<Dim, 4>(a);
    break;
  }
}

} // namespace

int ensure_moment_array(picnix_arena* a)
{
  if (a->d.um != nullptr)
    return PICNIX_OK;
  const size_t bytes = (size_t)a->g.nchunk * a->g.Ng * a->g.Ns * NMOM * sizeof(double);
  PICNIX_CUDA(a, cudaMalloc((void**)&a->d.um, bytes));
  PICNIX_CUDA(a, cudaMemsetAsync(a->d.um, 0, bytes, a->stream));
  return PICNIX_OK;
}

int launch_deposit_moment(picnix_arena* a)
{
  if (!a->particles_allocated)
    return fail(a, PICNIX_ERR_INVALID, "no particles allocated");
  int status = ensure_moment_array(a);
  if (status != PICNIX_OK)
    return status;
  const Geom& g = a->g;
  // the row-tile kernel reads through a pending lazy-sort permutation; everything else wants ordered arrays
  const int  npts     = g.dimension == 3 ? (g.order + 1) * (g.order + 1) * (g.order + 1)
                                          : (g.dimension == 2 ? (g.order + 1) * (g.order + 1) : g.order + 1);
  const bool row_path = a->pindex_valid && g.order % 2 == 0 && npts <= 32 && !a->force_generic;
  if (!row_path && (status = materialize_sort(a)) != PICNIX_OK)
    return status;
  // fill_all(um, 0), pic/engine/moment.hpp:104,166
  PICNIX_CUDA(a, cudaMemsetAsync(a->d.um, 0, (size_t)g.nchunk * g.Ng * g.Ns * NMOM * sizeof(double),
                                 a->stream));
  switch (g.dimension) {
  case 1:
    launch_moment_order<1>(a);
    break;
  case 2:
    launch_moment_order<2>(a);
    break;
  default:
    launch_moment_order<3>(a);
    break;
  }
  return check_cuda(a, cudaGetLastError(), "deposit_moment");
}

int launch_moment_halo_local(picnix_arena* a)
{
  int status = ensure_moment_array(a);
  if (status != PICNIX_OK)
    return status;
  const int64_t n = (int64_t)a->g.nchunk * a->g.Ng * a->g.Ns * NMOM;
  moment_halo_local_kernel<<<(unsigned)((n + 255) / 256), 256, 0, a->stream>>>(a->g, a->d);
  a->kernel_launches++;
  return check_cuda(a, cudaGetLastError(), "moment halo");
}

int launch_particle_energy(picnix_arena* a, double* particle)
{
  int status = ensure_moment_array(a);
  if (status != PICNIX_OK)
    return status;
  const int n = a->g.nchunk * a->g.Ns;
  double*   d_out = nullptr;
  PICNIX_CUDA(a, cudaMalloc((void**)&d_out, n * sizeof(double)));
  particle_energy_kernel<<<n, 256, 0, a->stream>>>(a->g, a->d, d_out);
  a->kernel_launches++;
  cudaError_t err = cudaMemcpyAsync(particle, d_out, n * sizeof(double), cudaMemcpyDeviceToHost, a->stream);
  if (err == cudaSuccess)
    err = cudaStreamSynchronize(a->stream);
  cudaFree(d_out);
  return check_cuda(a, err, "particle energy");
}

} // namespace picnix
