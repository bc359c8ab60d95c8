// -*- C++ -*-
// Arena life cycle and state transfer for libpicnix_b200.so.
//
// The arena is the device-resident equivalent of "all PicChunk objects of one MPI rank":
// geometry follows nix::Chunk::set_index_bounds / set_coordinate (nix/chunk.cpp:118-237),
// neighbour tables follow nix::ChunkVector::set_neighbors (nix/chunkvector.hpp:56-81) and the
// particle containers follow nix::XtensorParticle::allocate (nix/xtensor_particle.hpp:49-65).
#include "arena.hpp"
#include "transpose_kernels.cuh"

#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace picnix
{

int fail(picnix_arena* a, int code, const std::string& msg)
{
  if (a != nullptr)
    a->error = msg;
  return code;
}

int check_cuda(picnix_arena* a, cudaError_t err, const char* what)
{
  if (err == cudaSuccess)
    return PICNIX_OK;
  std::string msg = std::string(what) + ": " + cudaGetErrorString(err);
  return fail(a, PICNIX_ERR_CUDA, msg);
}

template <typename T>
static int dev_alloc(picnix_arena* a, T** ptr, size_t count, bool zero = true)
{
  size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  PICNIX_CUDA(a, cudaMalloc((void**)ptr, bytes));
  if (zero)
    PICNIX_CUDA(a, cudaMemset(*ptr, 0, bytes));
  return PICNIX_OK;
}

template <typename T>
static void dev_free(T*& ptr)
{
  if (ptr != nullptr) {
    cudaFree(ptr);
    ptr = nullptr;
  }
}

static int ensure_stage(picnix_arena* a, int64_t elems)
{
  if (elems <= a->stage_elems)
    return PICNIX_OK;
  if (a->h_stage)
    cudaFreeHost(a->h_stage);
  dev_free(a->d_stage);
  a->stage_elems = 0;
  PICNIX_CUDA(a, cudaMallocHost((void**)&a->h_stage, elems * sizeof(double)));
  PICNIX_CUDA(a, cudaMalloc((void**)&a->d_stage, elems * sizeof(double)));
  a->stage_elems = elems;
  return PICNIX_OK;
}

int upload_particles(picnix_arena* a, int ichunk, int is, const double* aos, int np)
{
  if (!a->particles_allocated)
    return fail(a, PICNIX_ERR_INVALID, "set_particle_capacity must precede upload_particles");
  if (ichunk < 0 || ichunk >= a->g.nchunk || is < 0 || is >= a->g.Ns)
    return fail(a, PICNIX_ERR_INVALID, "upload_particles: bad chunk/species index");
  int seg = ichunk * a->g.Ns + is;
  if (np < 0 || np > a->seg_cap[seg])
    return fail(a, PICNIX_ERR_OVERFLOW, "upload_particles: np exceeds segment capacity");

  {
    int mstatus = materialize_sort(a);
    if (mstatus != PICNIX_OK)
      return mstatus;
  }
  a->pindex_valid = false; // the new particles are not cell-ordered until the next sort
  a->leave_list_valid = false;
  int64_t elems = (int64_t)np * NC;
  if (np > 0) {
    int status = ensure_stage(a, elems);
    if (status != PICNIX_OK)
      return status;
    // make sure the previous use of the staging buffers has drained
    PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
    std::memcpy(a->h_stage, aos, elems * sizeof(double));
    PICNIX_CUDA(a, cudaMemcpyAsync(a->d_stage, a->h_stage, elems * sizeof(double),
                                   cudaMemcpyHostToDevice, a->stream));
    int threads = 256;
    int blocks  = (int)((elems + threads - 1) / threads);
    aos_to_soa_kernel<<<blocks, threads, 0, a->stream>>>(a->d_stage, a->d.xu, a->seg_off[seg],
                                                         a->d.pcap, np);
    a->kernel_launches++;
  }
  PICNIX_CUDA(a, cudaMemcpyAsync(a->d.np + seg, &np, sizeof(int), cudaMemcpyHostToDevice,
                                 a->stream));
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  return PICNIX_OK;
}

int download_particles(picnix_arena* a, int ichunk, int is, int which, int n, double* aos)
{
  if (!a->particles_allocated)
    return fail(a, PICNIX_ERR_INVALID, "no particles allocated");
  if (ichunk < 0 || ichunk >= a->g.nchunk || is < 0 || is >= a->g.Ns)
    return fail(a, PICNIX_ERR_INVALID, "download_particles: bad chunk/species index");
  int seg = ichunk * a->g.Ns + is;
  if (n < 0 || n > a->seg_cap[seg])
    return fail(a, PICNIX_ERR_INVALID, "download_particles: n exceeds segment capacity");
  if (n == 0)
    return PICNIX_OK;

  int64_t elems  = (int64_t)n * NC;
  int     status = materialize_sort(a);
  if (status != PICNIX_OK)
    return status;
  if ((status = ensure_stage(a, elems)) != PICNIX_OK)
    return status;
  const double* soa     = which == 0 ? a->d.xu : a->d.xv;
  int           threads = 256;
  int           blocks  = (int)((elems + threads - 1) / threads);
  soa_to_aos_kernel<<<blocks, threads, 0, a->stream>>>(soa, a->d_stage, a->seg_off[seg], a->d.pcap,
                                                       n);
  a->kernel_launches++;
  PICNIX_CUDA(a, cudaMemcpyAsync(a->h_stage, a->d_stage, elems * sizeof(double),
                                 cudaMemcpyDeviceToHost, a->stream));
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  std::memcpy(aos, a->h_stage, elems * sizeof(double));
  return PICNIX_OK;
}

static int field_info(picnix_arena* a, int ichunk, int which, double** dptr, int64_t* elems)
{
  if (ichunk < 0 || ichunk >= a->g.nchunk)
    return fail(a, PICNIX_ERR_INVALID, "bad chunk index");
  int64_t ncell = a->g.Ng;
  switch (which) {
  case PICNIX_FIELD_UF:
    *dptr  = a->d.uf + (int64_t)ichunk * ncell * 6;
    *elems = ncell * 6;
    return PICNIX_OK;
  case PICNIX_FIELD_UJ:
    *dptr  = a->d.uj + (int64_t)ichunk * ncell * 4;
    *elems = ncell * 4;
    return PICNIX_OK;
  case PICNIX_FIELD_FF:
    *dptr  = a->d.ff + (int64_t)ichunk * ncell * 9;
    *elems = ncell * 18; // host layout
    return PICNIX_OK;
  case PICNIX_FIELD_UM: {
    int status = ensure_moment_array(a);
    if (status != PICNIX_OK)
      return status;
    *dptr  = a->d.um + (int64_t)ichunk * ncell * a->g.Ns * 14;
    *elems = ncell * a->g.Ns * 14;
    return PICNIX_OK;
  }
  default:
    return fail(a, PICNIX_ERR_INVALID, "unsupported field selector");
  }
}

static int transfer_field(picnix_arena* a, int ichunk, int which, double* host, bool upload)
{
  double* dptr  = nullptr;
  int64_t elems = 0;
  int     status = field_info(a, ichunk, which, &dptr, &elems);
  if (status != PICNIX_OK)
    return status;
  int64_t ncell = a->g.Ng;
  if (which == PICNIX_FIELD_FF) {
    status = ensure_stage(a, elems);
    if (status != PICNIX_OK)
      return status;
    int threads = 256;
    PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
    if (upload) {
      PICNIX_CUDA(a, cudaMemcpyAsync(a->d_stage, host, elems * sizeof(double),
                                     cudaMemcpyHostToDevice, a->stream));
      int blocks = (int)((ncell * 9 + threads - 1) / threads);
      ff_compact_kernel<<<blocks, threads, 0, a->stream>>>(a->d_stage, dptr, ncell);
    } else {
      int blocks = (int)((ncell * 18 + threads - 1) / threads);
      ff_expand_kernel<<<blocks, threads, 0, a->stream>>>(dptr, a->d_stage, ncell);
      PICNIX_CUDA(a, cudaMemcpyAsync(host, a->d_stage, elems * sizeof(double),
                                     cudaMemcpyDeviceToHost, a->stream));
    }
    a->kernel_launches++;
  } else {
    if (upload) {
      PICNIX_CUDA(a, cudaMemcpyAsync(dptr, host, elems * sizeof(double), cudaMemcpyHostToDevice,
                                     a->stream));
    } else {
      PICNIX_CUDA(a, cudaMemcpyAsync(host, dptr, elems * sizeof(double), cudaMemcpyDeviceToHost,
                                     a->stream));
    }
  }
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  return PICNIX_OK;
}

static int build_geometry(picnix_arena* a)
{
  const picnix_config_t& c = a->cfg;
  Geom&                  g = a->g;

  if (c.order < 1 || c.order > 4)
    return fail(a, PICNIX_ERR_INVALID, "order must be 1..4");
  if (c.pusher < 0 || c.pusher > 2)
    return fail(a, PICNIX_ERR_INVALID, "Invalid pusher");
  if (c.interp < 0 || c.interp > 1)
    return fail(a, PICNIX_ERR_INVALID, "Invalid interpolation");
  if (c.Ns < 1 || c.nrank < 1 || c.rank < 0 || c.rank >= c.nrank)
    return fail(a, PICNIX_ERR_INVALID, "invalid Ns / rank configuration");

  g.order  = c.order;
  g.is_odd = c.order % 2;
  g.nb     = (c.order + 3) / 2; // pic/pic_chunk.cpp:195
  g.Ns     = c.Ns;
  g.cc     = c.cc;
  g.theta  = c.friedman;
  g.del[0] = c.delz;
  g.del[1] = c.dely;
  g.del[2] = c.delx;

  for (int i = 0; i < 3; i++) {
    if (c.ndims[i] < 1 || c.cdims[i] < 1 || c.ndims[i] % c.cdims[i] != 0)
      return fail(a, PICNIX_ERR_INVALID, "ndims must be a positive multiple of cdims");
    g.has_dim[i] = (c.ndims[i] == 1 && c.cdims[i] == 1) ? 0 : 1; // nix/application.cpp:262-266
    g.dims[i]    = c.ndims[i] / c.cdims[i];
    // nix/chunk.cpp:134-169
    g.Lb[i] = g.nb;
    g.Ub[i] = g.has_dim[i] ? g.nb + g.dims[i] - 1 : g.nb;
    g.M[i]  = g.dims[i] + 2 * g.nb;
    if (g.has_dim[i] && g.dims[i] < g.nb)
      return fail(a, PICNIX_ERR_INVALID,
                  "Number of grid points smaller than the minimum for the chosen shape order");
    g.glim[i][0] = 0.0;
    g.glim[i][1] = c.ndims[i] * g.del[i];
  }

  // pic/pic_chunk.cpp:9-23
  bool is_3d = g.has_dim[0] && g.has_dim[1] && g.has_dim[2];
  bool is_2d = !g.has_dim[0] && g.has_dim[1] && g.has_dim[2];
  bool is_1d = !g.has_dim[0] && !g.has_dim[1] && g.has_dim[2];
  if (!is_1d && !is_2d && !is_3d)
    return fail(a, PICNIX_ERR_INVALID, "Invalid dimension");
  g.dimension = is_3d ? 3 : (is_2d ? 2 : 1);

  g.Ng  = g.M[0] * g.M[1] * g.M[2];
  g.fsy = g.Ub[2] - g.Lb[2] + 2;
  g.fsz = g.fsy * (g.Ub[1] - g.Lb[1] + 2);
  return PICNIX_OK;
}

static int build_decomposition(picnix_arena* a, const int32_t* boundary)
{
  const picnix_config_t& c = a->cfg;
  Geom&                  g = a->g;
  const int              Cz = c.cdims[0], Cy = c.cdims[1], Cx = c.cdims[2];

  a->nchunk_global = Cz * Cy * Cx;
  if (c.nrank > a->nchunk_global)
    return fail(a, PICNIX_ERR_INVALID, "Number of processes should not exceed number of chunks");

  sfc_build(Cz, Cy, Cx, a->chunkid, a->coord);

  if (boundary != nullptr) {
    a->boundary.assign(boundary, boundary + c.nrank + 1);
    if (a->boundary.front() != 0 || a->boundary.back() != a->nchunk_global)
      return fail(a, PICNIX_ERR_INVALID, "rank boundary must span [0, nchunk]");
    for (int r = 0; r < c.nrank; r++)
      if (a->boundary[r + 1] < a->boundary[r])
        return fail(a, PICNIX_ERR_INVALID, "rank boundary must be ascending");
  } else {
    std::vector<double> load(a->nchunk_global, 1.0);
    a->boundary = assign_initial(load, c.nrank);
  }

  a->chunk_begin = a->boundary[c.rank];
  g.nchunk       = a->boundary[c.rank + 1] - a->boundary[c.rank];
  if (g.nchunk < 1)
    return fail(a, PICNIX_ERR_INVALID, "rank owns no chunk");

  auto rank_of = [&](int id) {
    // ChunkMap::get_rank, nix/chunkmap.cpp:109-116
    if (id < 0 || id >= a->nchunk_global)
      return -1;
    auto it = std::upper_bound(a->boundary.begin(), a->boundary.end(), id);
    return (int)(it - a->boundary.begin()) - 1;
  };
  auto neighbor_coord = [&](int coord, int delta, int dir) {
    // ChunkMap::get_neighbor_coord, nix/chunkmap.cpp:95-107
    int cdir = coord + delta;
    if (c.periodic[dir] == 1) {
      cdir = cdir >= 0 ? cdir : c.cdims[dir] - 1;
      cdir = cdir < c.cdims[dir] ? cdir : 0;
    } else {
      cdir = (cdir >= 0 && cdir < c.cdims[dir]) ? cdir : -1;
    }
    return cdir;
  };

  a->nbid.assign((size_t)g.nchunk * NBSIZE, -1);
  a->nbrank.assign((size_t)g.nchunk * NBSIZE, -1);
  a->nbr_code.assign((size_t)g.nchunk * NBSIZE, NB_NONE);
  std::vector<double> clim((size_t)g.nchunk * 6);

  for (int ic = 0; ic < g.nchunk; ic++) {
    int id     = a->chunk_begin + ic;
    int cc3[3] = {a->coord[3 * id + 2], a->coord[3 * id + 1], a->coord[3 * id + 0]}; // z,y,x

    for (int dz = -1; dz <= 1; dz++) {
      for (int dy = -1; dy <= 1; dy++) {
        for (int dx = -1; dx <= 1; dx++) {
          int k  = 9 * (dz + 1) + 3 * (dy + 1) + (dx + 1);
          int nz = neighbor_coord(cc3[0], dz, 0);
          int ny = neighbor_coord(cc3[1], dy, 1);
          int nx = neighbor_coord(cc3[2], dx, 2);
          int nb = -1;
          if (nz >= 0 && ny >= 0 && nx >= 0)
            nb = a->chunkid[(size_t)(nz * Cy + ny) * Cx + nx];
          a->nbid[(size_t)ic * NBSIZE + k]   = nb;
          a->nbrank[(size_t)ic * NBSIZE + k] = rank_of(nb);
        }
      }
    }

    // Chunk::set_coordinate with the offsets of nix/application.cpp:291-300
    for (int i = 0; i < 3; i++) {
      // = cc * ndims / cdims of the reference, without its 32-bit overflow for cc * ndims >= 2^31
      int    offset = cc3[i] * g.dims[i];
      double lo     = offset * g.del[i];
      double hi     = offset * g.del[i] + g.dims[i] * g.del[i];
      clim[(size_t)ic * 6 + 2 * i + 0] = lo;
      clim[(size_t)ic * 6 + 2 * i + 1] = hi;
    }
  }

  int status = dev_alloc(a, &a->d.clim, clim.size());
  if (status != PICNIX_OK)
    return status;
  PICNIX_CUDA(a, cudaMemcpy(a->d.clim, clim.data(), clim.size() * sizeof(double),
                            cudaMemcpyHostToDevice));
  return PICNIX_OK;
}

int build_comm_plan(picnix_arena* a); // halo.cu

} // namespace picnix

using namespace picnix;

extern "C" {

int picnix_cuda_arena_create(const picnix_config_t* cfg, const int32_t* boundary,
                             picnix_arena_t** out)
{
  if (cfg == nullptr || out == nullptr)
    return PICNIX_ERR_INVALID;
  *out = nullptr;

  int         ndev = 0;
  cudaError_t err  = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0) {
    // the product path has no CPU fallback: fail loudly
    std::fprintf(stderr, "[picnix_b200] no CUDA device available (%s); there is no CPU fallback\n",
                 err == cudaSuccess ? "device count is 0" : cudaGetErrorString(err));
    return PICNIX_ERR_NODEVICE;
  }

  auto a = new picnix_arena();
  // tuning/testing override of the row-kernel variant (same as set_option("deposit_mma"))
  if (const char* env = std::getenv("PICNIX_DEPOSIT_MMA"))
    a->deposit_mma = std::atoi(env) != 0;
  if (const char* env = std::getenv("PICNIX_ROW_KERNEL"))
    a->row_version = std::atoi(env) == 1 ? 1 : 2;
  if (const char* env = std::getenv("PICNIX_LAZY_SORT"))
    a->lazy_sort = std::atoi(env) != 0;
  a->cfg = *cfg;
  std::memset(&a->g, 0, sizeof(a->g));
  std::memset(&a->d, 0, sizeof(a->d));
  *out = a; // returned even on failure so that last_error can be read; caller destroys it

  int status = build_geometry(a);
  if (status != PICNIX_OK)
    return status;

  PICNIX_CUDA(a, cudaStreamCreateWithFlags(&a->stream, cudaStreamNonBlocking));
  a->own_stream = true;

  status = build_decomposition(a, boundary);
  if (status != PICNIX_OK)
    return status;

  const Geom& g     = a->g;
  int64_t     ncell = (int64_t)g.nchunk * g.Ng;
  a->nseg           = g.nchunk * g.Ns;

  if ((status = dev_alloc(a, &a->d.uf, ncell * 6)) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d.uj, ncell * 4)) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d.ff, ncell * 9)) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d.qm, g.Ns * 2)) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d.errflag, 4)) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d_reduce, (size_t)g.nchunk * 4)) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d.seg_stat, 4)) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d.spill_count, 1)) != PICNIX_OK)
    return status;
  PICNIX_CUDA(a, cudaMallocHost((void**)&a->h_stat, 4 * sizeof(int)));
  PICNIX_CUDA(a, cudaEventCreateWithFlags(&a->stat_event, cudaEventDisableTiming));
  if (const char* env = std::getenv("PICNIX_CHECK_GROWTH"))
    a->check_growth_always = std::atoi(env) != 0;
  a->d.far_cap = 1 << 18;
  if ((status = dev_alloc(a, &a->d.far_count, 1)) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d.far_rec, (size_t)a->d.far_cap * 8, false)) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d.np, a->nseg)) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d.ntail, a->nseg)) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d.pindex, (size_t)a->nseg * (g.Ng + 1))) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d.pcount, (size_t)a->nseg * (g.Ng + 1))) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d.seg_off, a->nseg)) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d.seg_cap, a->nseg)) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d.nbr, (size_t)g.nchunk * NBSIZE)) != PICNIX_OK)
    return status;

  a->seg_off.assign(a->nseg, 0);
  a->seg_cap.assign(a->nseg, 0);

  return build_comm_plan(a);
}

int picnix_cuda_arena_destroy(picnix_arena_t* a)
{
  if (a == nullptr)
    return PICNIX_OK;
  if (a->stream)
    cudaStreamSynchronize(a->stream);
  dev_free(a->d.uf);
  dev_free(a->d.uj);
  dev_free(a->d.ff);
  dev_free(a->d.um);
  dev_free(a->d.clim);
  dev_free(a->d.nbr);
  dev_free(a->d.xu);
  dev_free(a->d.xv);
  dev_free(a->d.gindex);
  dev_free(a->d.perm);
  dev_free(a->d.seg_off);
  dev_free(a->d.seg_cap);
  dev_free(a->d.np);
  dev_free(a->d.ntail);
  dev_free(a->d.pindex);
  dev_free(a->d.pcount);
  dev_free(a->d.qm);
  dev_free(a->d.errflag);
  dev_free(a->d.far_count);
  dev_free(a->d.far_rec);
  dev_free(a->d.leave_count);
  dev_free(a->d.leave_idx);
  dev_free(a->d.spill_count);
  dev_free(a->d.spill_rec);
  dev_free(a->d.seg_stat);
  if (a->h_stat)
    cudaFreeHost(a->h_stat);
  if (a->stat_event)
    cudaEventDestroy(a->stat_event);
  dev_free(a->d_reduce);
  hostio_destroy(a);
  dev_free(a->d_stage);
  dev_free(a->d_slot_peer);
  dev_free(a->d_slot_dst);
  dev_free(a->d_psend_ptrs);
  dev_free(a->d_psend_cnts);
  dev_free(a->d_psend_caps);
  if (a->d_scan_tmp)
    cudaFree(a->d_scan_tmp);
  if (a->h_stage)
    cudaFreeHost(a->h_stage);
  if (a->h_mig)
    cudaFreeHost(a->h_mig);
  if (a->h_bounds)
    cudaFreeHost(a->h_bounds);
  if (a->mig_event)
    cudaEventDestroy(a->mig_event);
  for (auto& p : a->peers) {
    dev_free(p.d_send_desc);
    dev_free(p.d_recv_desc);
    for (int m = 0; m < 3; m++) {
      dev_free(p.d_send_off[m]);
      dev_free(p.d_recv_off[m]);
      if (m == 2) { // Emf / Cur buffers are slices of d_send_all / d_recv_all
        dev_free(p.d_send[m]);
        dev_free(p.d_recv[m]);
      }
    }
    dev_free(p.d_psend);
    dev_free(p.d_precv);
    p.d_psend_count = nullptr; // slices of d_mig_counts
    p.d_rcount      = nullptr;
  }
  for (int m = 0; m < 2; m++) {
    dev_free(a->d_send_all[m]);
    dev_free(a->d_recv_all[m]);
    dev_free(a->d_send_off_all[m]);
    dev_free(a->d_recv_off_all[m]);
  }
  dev_free(a->d_mig_counts);
  dev_free(a->d_send_desc_all);
  dev_free(a->d_recv_desc_all);
  if (a->own_stream && a->stream)
    cudaStreamDestroy(a->stream);
  delete a;
  return PICNIX_OK;
}

const char* picnix_cuda_last_error(const picnix_arena_t* a)
{
  return a == nullptr ? "null arena" : a->error.c_str();
}

int picnix_cuda_set_stream(picnix_arena_t* a, void* stream)
{
  if (a == nullptr)
    return PICNIX_ERR_INVALID;
  cudaStreamSynchronize(a->stream);
  if (a->own_stream && a->stream)
    cudaStreamDestroy(a->stream);
  a->stream     = (cudaStream_t)stream;
  a->own_stream = false;
  return PICNIX_OK;
}

int picnix_cuda_synchronize(picnix_arena_t* a)
{
  if (a == nullptr)
    return PICNIX_ERR_INVALID;
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  // surface device-side error flags
  int flags[4];
  PICNIX_CUDA(a, cudaMemcpy(flags, a->d.errflag, sizeof(flags), cudaMemcpyDeviceToHost));
  if (flags[0] != 0)
    return fail(a, PICNIX_ERR_OVERFLOW,
                "particle spill list overflow: more migrants than segments and spill list can hold in one step");
  // flags[1] (a peer's message was full) is not an error any more: the records wait on the spill list and
  // leave with the next exchange (grow.cu; counted by picnix_cuda_get_growth_stats)
  if (flags[2] != 0)
    return fail(a, PICNIX_ERR_INVALID,
                "received a particle for a chunk/species this rank does not own (decomposition or "
                "message plan mismatch between ranks)");
  if (flags[3] != 0)
    return fail(a, PICNIX_ERR_OVERFLOW, "too many particles moved more than one cell in a step");
  return PICNIX_OK;
}

int picnix_cuda_get_layout(const picnix_arena_t* a, int32_t* nchunk, int32_t* chunk_id_begin,
                           int32_t* padded_dims, int32_t* margin, int32_t* Ng)
{
  if (a == nullptr)
    return PICNIX_ERR_INVALID;
  if (nchunk)
    *nchunk = a->g.nchunk;
  if (chunk_id_begin)
    *chunk_id_begin = a->chunk_begin;
  if (padded_dims) {
    padded_dims[0] = a->g.M[0];
    padded_dims[1] = a->g.M[1];
    padded_dims[2] = a->g.M[2];
  }
  if (margin)
    *margin = a->g.nb;
  if (Ng)
    *Ng = a->g.Ng;
  return PICNIX_OK;
}

int picnix_cuda_get_neighbors(const picnix_arena_t* a, int32_t ichunk, int32_t* nbid,
                              int32_t* nbrank)
{
  if (a == nullptr || ichunk < 0 || ichunk >= a->g.nchunk)
    return PICNIX_ERR_INVALID;
  for (int k = 0; k < NBSIZE; k++) {
    if (nbid)
      nbid[k] = a->nbid[(size_t)ichunk * NBSIZE + k];
    if (nbrank)
      nbrank[k] = a->nbrank[(size_t)ichunk * NBSIZE + k];
  }
  return PICNIX_OK;
}

int picnix_cuda_set_species(picnix_arena_t* a, int32_t is, double q, double m)
{
  if (a == nullptr || is < 0 || is >= a->g.Ns)
    return PICNIX_ERR_INVALID;
  double qm[2] = {q, m};
  PICNIX_CUDA(a, cudaMemcpy(a->d.qm + 2 * is, qm, sizeof(qm), cudaMemcpyHostToDevice));
  return PICNIX_OK;
}

int picnix_cuda_set_particle_capacity(picnix_arena_t* a, const int32_t* np_alloc)
{
  if (a == nullptr || np_alloc == nullptr)
    return PICNIX_ERR_INVALID;
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  dev_free(a->d.xu);
  dev_free(a->d.xv);
  dev_free(a->d.gindex);
  dev_free(a->d.perm);
  dev_free(a->d.leave_count);
  dev_free(a->d.leave_idx);
  dev_free(a->d.spill_rec);
  a->leave_list_valid = false;
  a->perm_pending     = false;
  a->stat_known       = false;
  a->stat_pending     = false;

  int64_t total = 0;
  for (int s = 0; s < a->nseg; s++) {
    if (np_alloc[s] < 0)
      return fail(a, PICNIX_ERR_INVALID, "negative particle capacity");
    // Particle::round_up_alloc, nix/particle.hpp:146-153
    int cap       = ((np_alloc[s] + ALLOC_UNIT) / ALLOC_UNIT) * ALLOC_UNIT;
    a->seg_off[s] = total;
    a->seg_cap[s] = cap;
    total += cap;
  }
  a->d.pcap = total;

  int status;
  if ((status = dev_alloc(a, &a->d.xu, (size_t)total * NC)) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d.xv, (size_t)total * NC)) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d.gindex, (size_t)total)) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d.perm, (size_t)total, false)) != PICNIX_OK)
    return status;
  // a quarter of the population leaving in one step is far beyond any Courant-limited run; the
  // migration falls back to scanning the keys if the list overflows
  a->d.leave_cap = (int)std::min<int64_t>(std::max<int64_t>(4096, total / 4), 1 << 30);
  if ((status = dev_alloc(a, &a->d.leave_count, 1)) != PICNIX_OK)
    return status;
  if ((status = dev_alloc(a, &a->d.leave_idx, (size_t)a->d.leave_cap, false)) != PICNIX_OK)
    return status;
  // migrants that do not fit wait here until their segment has grown (grow.cu); the list of particles
  // that moved more than a cell (rowpush.cu) scales with the population as well
  a->d.spill_cap = (int)std::min<int64_t>(std::max<int64_t>(65536, total / 16), 1 << 28);
  if ((status = dev_alloc(a, &a->d.spill_rec, (size_t)a->d.spill_cap * 8, false)) != PICNIX_OK)
    return status;
  PICNIX_CUDA(a, cudaMemset(a->d.spill_count, 0, sizeof(int)));
  {
    const int want = (int)std::min<int64_t>(std::max<int64_t>(1 << 18, total / 16), 1 << 28);
    if (want > a->d.far_cap) {
      dev_free(a->d.far_rec);
      a->d.far_cap = want;
      if ((status = dev_alloc(a, &a->d.far_rec, (size_t)want * 8, false)) != PICNIX_OK)
        return status;
    }
  }
  PICNIX_CUDA(a, cudaMemcpy(a->d.seg_off, a->seg_off.data(), a->nseg * sizeof(int64_t),
                            cudaMemcpyHostToDevice));
  PICNIX_CUDA(a, cudaMemcpy(a->d.seg_cap, a->seg_cap.data(), a->nseg * sizeof(int32_t),
                            cudaMemcpyHostToDevice));
  PICNIX_CUDA(a, cudaMemset(a->d.np, 0, a->nseg * sizeof(int)));
  PICNIX_CUDA(a, cudaMemset(a->d.ntail, 0, a->nseg * sizeof(int)));
  a->particles_allocated = true;
  return PICNIX_OK;
}

int picnix_cuda_upload_field(picnix_arena_t* a, int32_t ichunk, int32_t which, const double* host)
{
  if (a == nullptr || host == nullptr)
    return PICNIX_ERR_INVALID;
  return transfer_field(a, ichunk, which, const_cast<double*>(host), true);
}

int picnix_cuda_download_field(picnix_arena_t* a, int32_t ichunk, int32_t which, double* host)
{
  if (a == nullptr || host == nullptr)
    return PICNIX_ERR_INVALID;
  return transfer_field(a, ichunk, which, host, false);
}

int picnix_cuda_upload_particles(picnix_arena_t* a, int32_t ichunk, int32_t is,
                                 const double* xu_aos, int32_t np)
{
  if (a == nullptr || (xu_aos == nullptr && np > 0))
    return PICNIX_ERR_INVALID;
  return upload_particles(a, ichunk, is, xu_aos, np);
}

int picnix_cuda_download_particles(picnix_arena_t* a, int32_t ichunk, int32_t is, int32_t which,
                                   int32_t n, double* aos)
{
  if (a == nullptr || (aos == nullptr && n > 0))
    return PICNIX_ERR_INVALID;
  return download_particles(a, ichunk, is, which, n, aos);
}

int picnix_cuda_get_np(picnix_arena_t* a, int32_t* np)
{
  if (a == nullptr || np == nullptr)
    return PICNIX_ERR_INVALID;
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  PICNIX_CUDA(a, cudaMemcpy(np, a->d.np, a->nseg * sizeof(int), cudaMemcpyDeviceToHost));
  return PICNIX_OK;
}

int picnix_cuda_download_pindex(picnix_arena_t* a, int32_t ichunk, int32_t is, int32_t* pindex)
{
  if (a == nullptr || pindex == nullptr || ichunk < 0 || ichunk >= a->g.nchunk || is < 0 ||
      is >= a->g.Ns)
    return PICNIX_ERR_INVALID;
  int seg = ichunk * a->g.Ns + is;
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  PICNIX_CUDA(a, cudaMemcpy(pindex, a->d.pindex + (size_t)seg * (a->g.Ng + 1),
                            (a->g.Ng + 1) * sizeof(int), cudaMemcpyDeviceToHost));
  return PICNIX_OK;
}

int picnix_cuda_download_gindex(picnix_arena_t* a, int32_t ichunk, int32_t is, int32_t n,
                                int32_t* gindex)
{
  if (a == nullptr || gindex == nullptr || ichunk < 0 || ichunk >= a->g.nchunk || is < 0 ||
      is >= a->g.Ns || !a->particles_allocated)
    return PICNIX_ERR_INVALID;
  int seg = ichunk * a->g.Ns + is;
  if (n < 0 || n > a->seg_cap[seg])
    return PICNIX_ERR_INVALID;
  {
    int mstatus = materialize_sort(a);
    if (mstatus != PICNIX_OK)
      return mstatus;
  }
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  PICNIX_CUDA(a, cudaMemcpy(gindex, a->d.gindex + a->seg_off[seg], n * sizeof(int),
                            cudaMemcpyDeviceToHost));
  return PICNIX_OK;
}

int picnix_cuda_get_counters(const picnix_arena_t* a, int64_t* kernel_launches,
                             int64_t* particle_pushes)
{
  if (a == nullptr)
    return PICNIX_ERR_INVALID;
  if (kernel_launches)
    *kernel_launches = a->kernel_launches;
  if (particle_pushes)
    *particle_pushes = a->particle_pushes;
  return PICNIX_OK;
}

} // extern "C"
