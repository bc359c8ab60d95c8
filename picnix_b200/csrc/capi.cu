// -*- C++ -*-
// extern "C" entry points of the per-timestep phases (include/picnix_b200.h) and the whole-step
// schedule of PicApplication::push_openmp (pic/pic_application.cpp:219-292).
#include "arena.hpp"

using namespace picnix;

namespace
{
inline bool bad_range(const picnix_arena* a, int c0, int cn)
{
  if (cn < 0)
    return false;
  return c0 < 0 || c0 + cn > a->g.nchunk;
}
} // namespace

#define PICNIX_CHECK_RANGE(a, c0, cn)                                                              \
  if ((a) == nullptr)                                                                              \
    return PICNIX_ERR_INVALID;                                                                     \
  if (bad_range((a), (c0), (cn)))                                                                  \
    return fail((a), PICNIX_ERR_INVALID, "chunk range outside the arena");

extern "C" {

int picnix_cuda_set_option(picnix_arena_t* a, const char* key, int64_t value)
{
  if (a == nullptr || key == nullptr)
    return PICNIX_ERR_INVALID;
  if (std::string(key) == "force_generic") {
    a->force_generic = value != 0;
    return PICNIX_OK;
  }
  if (std::string(key) == "lazy_sort") {
    int status = materialize_sort(a);
    a->lazy_sort = value != 0;
    return status;
  }
  if (std::string(key) == "async_migration") {
    a->async_migration = value != 0;
    return PICNIX_OK;
  }
  if (std::string(key) == "check_growth") {
    a->check_growth_always = value != 0;
    return PICNIX_OK;
  }
  if (std::string(key) == "row_kernel") {
    if (value != 1 && value != 2)
      return fail(a, PICNIX_ERR_INVALID, "row_kernel must be 1 or 2");
    a->row_version = (int)value;
    return PICNIX_OK;
  }
  if (std::string(key) == "deposit_mma") {
    a->deposit_mma = value != 0;
    return PICNIX_OK;
  }
  return fail(a, PICNIX_ERR_INVALID, std::string("unknown option: ") + key);
}

int picnix_cuda_init_friedman(picnix_arena_t* a, int32_t c0, int32_t cn)
{
  PICNIX_CHECK_RANGE(a, c0, cn);
  return launch_init_friedman(a, c0, cn);
}

int picnix_cuda_push_bfd(picnix_arena_t* a, int32_t c0, int32_t cn, double delt)
{
  PICNIX_CHECK_RANGE(a, c0, cn);
  return launch_push_bfd(a, c0, cn, delt);
}

int picnix_cuda_push_efd(picnix_arena_t* a, int32_t c0, int32_t cn, double delt)
{
  PICNIX_CHECK_RANGE(a, c0, cn);
  return launch_push_efd(a, c0, cn, delt);
}

int picnix_cuda_push_velocity(picnix_arena_t* a, int32_t c0, int32_t cn, double delt)
{
  PICNIX_CHECK_RANGE(a, c0, cn);
  return launch_push_velocity(a, c0, cn, delt);
}

int picnix_cuda_push_position(picnix_arena_t* a, int32_t c0, int32_t cn, double delt)
{
  PICNIX_CHECK_RANGE(a, c0, cn);
  return launch_push_position(a, c0, cn, delt);
}

int picnix_cuda_deposit_current(picnix_arena_t* a, int32_t c0, int32_t cn, double delt)
{
  PICNIX_CHECK_RANGE(a, c0, cn);
  return launch_deposit_current(a, c0, cn, delt);
}

int picnix_cuda_deposit_moment(picnix_arena_t* a)
{
  if (a == nullptr)
    return PICNIX_ERR_INVALID;
  return launch_deposit_moment(a);
}

int picnix_cuda_get_particle_energy(picnix_arena_t* a, double* particle)
{
  if (a == nullptr || particle == nullptr)
    return PICNIX_ERR_INVALID;
  return launch_particle_energy(a, particle);
}

int picnix_cuda_sort_particle(picnix_arena_t* a, int32_t c0, int32_t cn)
{
  PICNIX_CHECK_RANGE(a, c0, cn);
  // PicChunk::sort_particle = count(0, Np-1, true, order) + sort()  (pic/pic_chunk.cpp:447-453)
  int status = launch_count(a, c0, cn);
  if (status != PICNIX_OK)
    return status;
  return launch_sort(a, c0, cn);
}

int picnix_cuda_push_deposit_fused(picnix_arena_t* a, int32_t c0, int32_t cn, double delt)
{
  PICNIX_CHECK_RANGE(a, c0, cn);
  return launch_push_deposit_fused(a, c0, cn, delt);
}

int picnix_cuda_boundary_begin(picnix_arena_t* a, int32_t mode)
{
  if (a == nullptr)
    return PICNIX_ERR_INVALID;
  return launch_halo_begin(a, mode);
}

int picnix_cuda_boundary_end(picnix_arena_t* a, int32_t mode)
{
  if (a == nullptr)
    return PICNIX_ERR_INVALID;
  return launch_halo_end(a, mode);
}

int picnix_cuda_step(picnix_arena_t* a, double delt, int32_t nstep)
{
  if (a == nullptr)
    return PICNIX_ERR_INVALID;
  if (a->cfg.nrank != 1)
    return fail(a, PICNIX_ERR_INVALID,
                "picnix_cuda_step needs nrank == 1; multi-rank callers drive the phases and move "
                "the peer buffers between boundary_begin and boundary_end");

#define STEP_CALL(expr)                                                                            \
  do {                                                                                             \
    int status_ = (expr);                                                                          \
    if (status_ != PICNIX_OK)                                                                      \
      return status_;                                                                              \
  } while (0)

  for (int step = 0; step < nstep; step++) {
    // phase A of push_openmp: B half step, particle push, deposit, start J and particle exchange
    STEP_CALL(launch_push_bfd(a, 0, -1, 0.5 * delt));
    STEP_CALL(launch_push_deposit_fused(a, 0, -1, delt));
    STEP_CALL(launch_halo_begin(a, PICNIX_BOUNDARY_CUR));
    STEP_CALL(launch_halo_begin(a, PICNIX_BOUNDARY_PARTICLE));
    STEP_CALL(launch_push_bfd(a, 0, -1, 0.5 * delt));
    // phase B: finish J exchange, E full step, start E/B exchange
    STEP_CALL(launch_halo_end(a, PICNIX_BOUNDARY_CUR));
    STEP_CALL(launch_push_efd(a, 0, -1, delt));
    STEP_CALL(launch_halo_begin(a, PICNIX_BOUNDARY_EMF));
    // phases C-E: particles arrive -> wrap, count, sort; fields arrive
    STEP_CALL(launch_halo_end(a, PICNIX_BOUNDARY_PARTICLE));
    STEP_CALL(launch_halo_end(a, PICNIX_BOUNDARY_EMF));
  }
#undef STEP_CALL
  return PICNIX_OK;
}

int picnix_cuda_get_diverror(picnix_arena_t* a, double* efd, double* bfd)
{
  if (a == nullptr || efd == nullptr || bfd == nullptr)
    return PICNIX_ERR_INVALID;
  return launch_diverror(a, efd, bfd);
}

int picnix_cuda_get_field_energy(picnix_arena_t* a, double* efd, double* bfd)
{
  if (a == nullptr || efd == nullptr || bfd == nullptr)
    return PICNIX_ERR_INVALID;
  return launch_field_energy(a, efd, bfd);
}

int picnix_cuda_step_host(picnix_arena_t* a, double delt, int32_t nstep, double* uf, double* uj,
                          double* ff, double* xu, const int32_t* np_in, const int32_t* np_cap,
                          int32_t* np_out)
{
  if (a == nullptr || uf == nullptr || uj == nullptr || ff == nullptr || xu == nullptr ||
      np_in == nullptr || np_cap == nullptr || np_out == nullptr)
    return PICNIX_ERR_INVALID;

  if (nstep < 0)
    return fail(a, PICNIX_ERR_INVALID, "nstep must be >= 0");
  if (a->cfg.nrank != 1)
    return fail(a, PICNIX_ERR_INVALID, "picnix_cuda_step_host needs nrank == 1");
  // pipelined transfer + step, hostio.cu
  return step_host_pipelined(a, delt, nstep, uf, uj, ff, xu, np_in, np_cap, np_out);
}

} // extern "C"
