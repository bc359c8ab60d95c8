/*
 * picnix_b200.h -- C ABI of the B200-native PIC-NIX hot path.
 *
 * This is the drop-in boundary: everything a `PicChunk` subclass of the reference
 * (amanotk/pic-nix, pic/pic_chunk.hpp:90-143) needs in order to run its per-timestep kernels on a
 * B200 instead of on the host.  Plain pointers and sizes only; every function returns a status
 * (PICNIX_OK == 0) instead of the reference's `ERROR << ...; MPI_Abort()` convention
 * (pic/pic_chunk.cpp:14-17), and `picnix_cuda_last_error()` returns the message the reference
 * would have logged.
 *
 * Data model.  One *arena* per GPU/rank holds all chunks the rank owns -- a contiguous range of
 * space-filling-curve chunk ids, exactly as nix::Application::setup_chunks_init assigns them
 * (nix/application.cpp:287-292) -- so that each kernel is launched once over all of them:
 *
 *   uf[chunk][Mz][My][Mx][6]   E,B            (pic/pic_chunk.cpp:114)   M* = dims + 2*margin
 *   uj[chunk][Mz][My][Mx][4]   rho,Jx,Jy,Jz   (pic/pic_chunk.cpp:115)
 *   ff[chunk][Mz][My][Mx][3][3] Friedman-filter history of E (the reference allocates [3][6] but
 *                              only uses [..][0:3], pic/engine/maxwell.hpp:44-61)
 *   particles: structure-of-arrays, 7 x f64 per particle (x,y,z,ux,uy,uz,id-bits;
 *              nix/particle.hpp:18) in two buffers `xu`/`xv` like nix::XtensorParticle
 *              (nix/xtensor_particle.hpp:15-19); one segment per (chunk, species)
 *   gindex[particle] i32 cell key, pindex[chunk][species][Ng+1] i32 first particle of each cell
 *
 * Host arrays crossing this boundary use the REFERENCE's layouts (AoS [Np][7] particles,
 * [..][3][6] filter array) so a maintainer can pass `xt::xtensor::data()` pointers directly.
 *
 * Threading: calls on one arena must not overlap; different arenas are independent.
 * All work is enqueued on the arena's CUDA stream; functions that return data synchronise it.
 */
#ifndef PICNIX_B200_H
#define PICNIX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes */
#define PICNIX_OK 0
#define PICNIX_ERR_INVALID 1   /* bad argument / unsupported configuration          */
#define PICNIX_ERR_CUDA 2      /* CUDA runtime error (message in last_error)        */
#define PICNIX_ERR_OVERFLOW 3  /* a particle segment or migration buffer overflowed */
#define PICNIX_ERR_NODEVICE 4  /* no CUDA device: there is NO CPU fallback          */

/* boundary-exchange modes, pic/pic.hpp:40-46 */
#define PICNIX_BOUNDARY_EMF 0
#define PICNIX_BOUNDARY_CUR 1
#define PICNIX_BOUNDARY_MOM 2
#define PICNIX_BOUNDARY_PARTICLE 3

/* field selectors for upload/download */
#define PICNIX_FIELD_UF 0 /* [Mz][My][Mx][6]        */
#define PICNIX_FIELD_UJ 1 /* [Mz][My][Mx][4]        */
#define PICNIX_FIELD_FF 2 /* [Mz][My][Mx][3][6]     */
#define PICNIX_FIELD_UM 3 /* [Mz][My][Mx][Ns][14]   */

/* physical boundary conditions on the faces of a non-periodic direction (picnix_cuda_set_boundary_condition) */
#define PICNIX_BC_NONE 0       /* open: fields as the halo exchange left them, particles leave          */
#define PICNIX_BC_CONDUCTING 1 /* conducting wall, specular reflection (example/mrx/main.cpp:183-382)   */
#define PICNIX_BC_WALL 2       /* shock-tube wall, momentum reversed (example/shock/main.cpp:232-283)   */
#define PICNIX_BC_INFLOW 3     /* upstream values imposed (example/shock/main.cpp:285-340)              */

/* pusher / interpolation enums, pic/engine/velocity.hpp:12-28 */
#define PICNIX_PUSHER_BORIS 0
#define PICNIX_PUSHER_VAY 1
#define PICNIX_PUSHER_HIGUERA_CARY 2
#define PICNIX_INTERP_MC 0
#define PICNIX_INTERP_WT 1

/*
 * Run configuration: the subset of config.toml `parameter` / `application.option`
 * (nix/cfgparser.hpp:111-217, pic/pic_chunk.cpp:135-262) that the hot path reads.
 */
typedef struct picnix_config {
  int32_t ndims[3];     /* global cells Nz,Ny,Nx                                      */
  int32_t cdims[3];     /* number of chunks Cz,Cy,Cx                                  */
  int32_t periodic[3];  /* ChunkMap periodicity z,y,x (nix/chunkmap.cpp:88-107)       */
  int32_t order;        /* shape-function order 1..4 (pic/pic_engine.hpp:20-22)       */
  int32_t pusher;       /* PICNIX_PUSHER_*                                            */
  int32_t interp;       /* PICNIX_INTERP_*                                            */
  int32_t Ns;           /* number of species                                          */
  int32_t nrank;        /* number of ranks (GPUs) the chunk ids are split over        */
  int32_t rank;         /* this arena's rank                                          */
  double  cc;           /* speed of light                                             */
  double  delx, dely, delz; /* cell sizes (Chunk::set_coordinate, nix/chunk.cpp:210)  */
  double  friedman;     /* Friedman filter theta (pic/engine/maxwell.hpp:41)          */
  double  buffer_ratio; /* particle buffer slack, default 0.2 (pic/pic_chunk.cpp:260) */
} picnix_config_t;

typedef struct picnix_arena picnix_arena_t;

/* ---- decomposition (host integer logic; bit-exact with the reference) ------------------------ */

/* Generalized Hilbert curve chunk ordering, nix/sfc.cpp:78-141.
 * chunkid: [Cz][Cy][Cx] -> id, coord: [Cz*Cy*Cx][3] -> (x,y,z) as stored by nix::ChunkMap. */
int picnix_sfc_build(int32_t Cz, int32_t Cy, int32_t Cx, int32_t* chunkid, int32_t* coord);

/* Balancer::assign_initial for a given per-chunk load, nix/balancer.cpp:71-124.
 * boundary has nrank+1 entries. */
int picnix_assign_initial(const double* load, int32_t nchunk, int32_t nrank, int32_t* boundary);

/* Balancer::assign (one SMILEI-style boundary adjustment), nix/balancer.cpp:8-69. */
int picnix_assign_rebalance(const double* load, int32_t nchunk, int32_t nrank, int32_t* boundary);

/* ---- arena life cycle ------------------------------------------------------------------------ */

/* `boundary` (nrank+1 ascending chunk ids) is ChunkMap::set_rank_boundary's argument
 * (nix/chunkmap.cpp:118-121); NULL means an even split by assign_initial with unit loads. */
int picnix_cuda_arena_create(const picnix_config_t* cfg, const int32_t* boundary,
                             picnix_arena_t** arena);
int picnix_cuda_arena_destroy(picnix_arena_t* arena);
const char* picnix_cuda_last_error(const picnix_arena_t* arena);

/* tuning / testing switches:
 *   "force_generic" = 1  bypasses the tiled kernels (3-D, 2-D, 1-D row-owner deposit): the thread-per-particle
 *                        kernels run instead
 *   "lazy_sort"     = 0  the counting sort always moves the particles (default 1: when the tiled
 *                        fused kernel will consume the result only the permutation is written and
 *                        the reordering rides on the next push; results are identical)
 *   "deposit_mma"   = 1  FP64-MMA formulation of the deposit (slower; kept as measured evidence)
 *   "row_kernel"    = 1  the round-1 tiled kernel (experiments/rowfused_v1.cu) instead of rowpush.cu (default 2)
 *   "check_growth"  = 1  look at the segment populations on the host before EVERY sort (default: only
 *                        when the previous step's statistics say a segment may fill up; see
 *                        picnix_cuda_get_growth_stats)
 *   "async_migration" = 1  multi-rank particle exchange without a host synchronisation: message
 *                        sizes follow from the previous step's counts (both sides compute the same
 *                        bound, get_comm_buffer returns send AND receive sizes, set_recv_bytes is not
 *                        needed), the actual count travels in a 64-byte header behind the records */
int picnix_cuda_set_option(picnix_arena_t* arena, const char* key, int64_t value);

/*
 * Physical boundary condition of one face of the global domain: axis 0 = z, 1 = y, 2 = x; side 0 = lower,
 * 1 = upper.  Replaces the set_boundary_field / set_boundary_particle hooks a problem's MainChunk
 * overrides in the reference (pic/pic_chunk.hpp:122-126): the fields are treated after every field halo
 * exchange, the particles right after the position push (before the cell count and the deposit).
 * PICNIX_BC_CONDUCTING is available for walls normal to y, PICNIX_BC_WALL for the lower and
 * PICNIX_BC_INFLOW for the upper x boundary (values = Ex, Ey, Ez, Bx, By, Bz imposed in the margin) --
 * the kinds and faces the reference's examples use.  The direction must be non-periodic in the arena's
 * configuration.
 */
int picnix_cuda_set_boundary_condition(picnix_arena_t* arena, int32_t axis, int32_t side, int32_t kind,
                                       const double* values /* [6] or NULL */);

/* PicChunk::inject_particle hook (pic/pic_chunk.hpp:128): append `n` host-generated particles (AoS [n][7])
 * to (chunk, species); call between boundary_begin and boundary_end of PICNIX_BOUNDARY_PARTICLE, whose
 * sort takes them in.  Segments grow as needed (picnix_cuda_get_growth_stats). */
int picnix_cuda_inject_particles(picnix_arena_t* arena, int32_t ichunk, int32_t is, const double* xu_aos,
                                 int32_t n);

/* use an existing CUDA stream (cudaStream_t as void*); default is a stream the arena owns */
int picnix_cuda_set_stream(picnix_arena_t* arena, void* stream);
int picnix_cuda_synchronize(picnix_arena_t* arena);

/* number of local chunks, first global chunk id, padded dims {Mz,My,Mx}, margin, Ng */
int picnix_cuda_get_layout(const picnix_arena_t* arena, int32_t* nchunk, int32_t* chunk_id_begin,
                           int32_t* padded_dims, int32_t* margin, int32_t* Ng);
/* nbid/nbrank[27] of a local chunk, index 9*(dz+1)+3*(dy+1)+(dx+1) (nix/chunk.hpp:176-197) */
int picnix_cuda_get_neighbors(const picnix_arena_t* arena, int32_t ichunk, int32_t* nbid,
                              int32_t* nbrank);

/* ---- state transfer (PicChunk::setup / get_internal_data / pack, pic/pic_chunk.cpp:59-122) --- */

int picnix_cuda_set_species(picnix_arena_t* arena, int32_t is, double q, double m);
/* np_alloc[nchunk*Ns]: requested capacity of each (chunk, species) segment; rounded up like
 * Particle::round_up_alloc (nix/particle.hpp:146-153).  Must precede upload_particles. */
int picnix_cuda_set_particle_capacity(picnix_arena_t* arena, const int32_t* np_alloc);
int picnix_cuda_upload_field(picnix_arena_t* arena, int32_t ichunk, int32_t which,
                             const double* host);
int picnix_cuda_download_field(picnix_arena_t* arena, int32_t ichunk, int32_t which, double* host);
/* xu_aos: [np][7] f64, the reference's XtensorParticle::xu layout */
int picnix_cuda_upload_particles(picnix_arena_t* arena, int32_t ichunk, int32_t is,
                                 const double* xu_aos, int32_t np);
/* which: 0 = xu, 1 = xv; copies the first n particles back into AoS [n][7] */
int picnix_cuda_download_particles(picnix_arena_t* arena, int32_t ichunk, int32_t is,
                                   int32_t which, int32_t n, double* aos);
int picnix_cuda_get_np(picnix_arena_t* arena, int32_t* np /* [nchunk*Ns] */);
int picnix_cuda_download_pindex(picnix_arena_t* arena, int32_t ichunk, int32_t is,
                                int32_t* pindex /* [Ng+1] */);
int picnix_cuda_download_gindex(picnix_arena_t* arena, int32_t ichunk, int32_t is, int32_t n,
                                int32_t* gindex);

/* ---- kernel entry points: one per PicChunk virtual (pic/pic_chunk.hpp:108-142) --------------- */
/* Each acts on local chunks [chunk_begin, chunk_begin+chunk_count); chunk_count < 0 = all.     */

int picnix_cuda_init_friedman(picnix_arena_t* arena, int32_t chunk_begin, int32_t chunk_count);
int picnix_cuda_push_bfd(picnix_arena_t* arena, int32_t chunk_begin, int32_t chunk_count,
                         double delt);
int picnix_cuda_push_efd(picnix_arena_t* arena, int32_t chunk_begin, int32_t chunk_count,
                         double delt);
int picnix_cuda_push_velocity(picnix_arena_t* arena, int32_t chunk_begin, int32_t chunk_count,
                              double delt);
/* includes XtensorParticle::count(0, Np-1, true, order) like pic_engine::Position::set_boundary
 * (pic/pic_engine.hpp:292-303) */
int picnix_cuda_push_position(picnix_arena_t* arena, int32_t chunk_begin, int32_t chunk_count,
                              double delt);
int picnix_cuda_deposit_current(picnix_arena_t* arena, int32_t chunk_begin, int32_t chunk_count,
                                double delt);
/* PicChunk::deposit_moment (pic/pic_chunk.cpp:525-534, pic/engine/moment.hpp): um <- 0, then the 14
 * velocity moments of every species on the (order+1)^dim momentum-conserving stencil.  All local
 * chunks; follow with boundary_begin/end(PICNIX_BOUNDARY_MOM) like PicApplication's diagnostics do. */
int picnix_cuda_deposit_moment(picnix_arena_t* arena);
/* PicChunk::sort_particle: count(reset) + counting sort, drops out-of-chunk particles */
int picnix_cuda_sort_particle(picnix_arena_t* arena, int32_t chunk_begin, int32_t chunk_count);

/* Fused K1+K2: push_velocity + push_position + count + deposit_current in ONE pass over the
 * particles (results equal the three separate calls; `xv` is not materialised). */
int picnix_cuda_push_deposit_fused(picnix_arena_t* arena, int32_t chunk_begin,
                                   int32_t chunk_count, double delt);

/* ---- boundary exchange (Chunk::{pack,begin,end,unpack}_bc_exchange, nix/chunk.hpp:392-543) --- */
/*
 * begin = set_boundary_pack + set_boundary_begin for ALL local chunks: neighbours inside the arena
 * are served directly by a device kernel (no messages); data for chunks owned by another rank is
 * packed into one contiguous device send buffer per peer rank.
 * end   = set_boundary_end + set_boundary_unpack: consumes the per-peer receive buffers.
 * For PICNIX_BOUNDARY_PARTICLE `end` also wraps, counts and sorts, as
 * XtensorHaloParticle3D::post_unpack does (nix/xtensor_halo3d.hpp:477-498).
 * Between begin and end the caller moves send -> recv buffers between ranks (NCCL / P2P / MPI).
 */
int picnix_cuda_boundary_begin(picnix_arena_t* arena, int32_t mode);
int picnix_cuda_boundary_end(picnix_arena_t* arena, int32_t mode);
/* number of peer ranks this arena exchanges halos with, and their ranks */
int picnix_cuda_get_peers(const picnix_arena_t* arena, int32_t* npeer, int32_t* peer_rank);
/* device pointers + byte counts of the send/recv buffer for `peer_index` after begin(mode).
 * For the particle mode send_bytes is exact after begin; the receiver learns its size from the
 * peer (first exchange the 8-byte counts, then call set_recv_bytes, then move the payload). */
int picnix_cuda_get_comm_buffer(picnix_arena_t* arena, int32_t mode, int32_t peer_index,
                                void** send_ptr, int64_t* send_bytes, void** recv_ptr,
                                int64_t* recv_bytes);
int picnix_cuda_set_recv_bytes(picnix_arena_t* arena, int32_t mode, int32_t peer_index,
                               int64_t recv_bytes);

/* ---- whole step (PicApplication::push_openmp, pic/pic_application.cpp:219-292) --------------- */
/* Single-rank arenas only (nrank == 1): runs nstep full time steps on the device. */
int picnix_cuda_step(picnix_arena_t* arena, double delt, int32_t nstep);

/* ---- diagnostics (PicChunk::get_diverror/get_energy, pic/pic_chunk.cpp:407-445) -------------- */
int picnix_cuda_get_diverror(picnix_arena_t* arena, double* efd, double* bfd /* [nchunk] each */);
int picnix_cuda_get_field_energy(picnix_arena_t* arena, double* efd, double* bfd);
/* particle part of PicChunk::get_energy (pic/pic_chunk.cpp:428-438): per chunk and species the sum
 * over interior cells of um[..][4]*c - um[..][0]*c^2 (rest mass subtracted); particle[nchunk*Ns].
 * Needs deposit_moment + the BoundaryMom exchange first, as in the reference. */
int picnix_cuda_get_particle_energy(picnix_arena_t* arena, double* particle);
/* Growing particle storage (XtensorParticle::resize, nix/xtensor_particle.hpp:70-115, as called by
 * XtensorHaloParticle3D::pre_unpack, nix/xtensor_halo3d.hpp:406-418).  A (chunk, species) segment that
 * fills up is enlarged before the particle exchange's sort; migrants that found it full wait on a spill
 * list and are appended afterwards, so a run never aborts and never loses particles because a chunk's
 * population grew.  segment_regrows: how often the particle arrays were re-laid out.  late_particles:
 * migrants that were appended (or sent to a peer) one step late because the fill-up was not foreseen by
 * the previous step's statistics; 0 in any Courant-limited run. */
int picnix_cuda_get_growth_stats(const picnix_arena_t* arena, int64_t* segment_regrows,
                                 int64_t* late_particles);
/* counters since arena creation: kernels launched by this library, particles pushed */
int picnix_cuda_get_counters(const picnix_arena_t* arena, int64_t* kernel_launches,
                             int64_t* particle_pushes);

/* ---- host-buffer convenience (what a host-resident PicChunk would call every step) ----------- */
/*
 * Upload the state of all local chunks from HOST arrays in the reference's layouts, run `nstep`
 * steps, and download the state back.  uf/uj/ff: [nchunk][...] concatenated; xu: AoS particles of
 * all (chunk, species) segments concatenated with `np_in[seg]` entries each and room for
 * `np_cap[seg]`; np_out receives the new counts.
 *
 * The call is a three-stream pipeline (copy-in / compute / copy-out, hostio.cu) and is bound by the
 * PCIe link.  The buffers are used in place: pageable buffers are page-locked on first use
 * (cudaHostRegister) and stay registered until the arena is destroyed, so pass the SAME arrays
 * every step, or allocate them with picnix_cuda_host_alloc.  `uj` is an output only when
 * nstep >= 1 (the deposit starts from zero, pic/engine/current.hpp:91).
 */
int picnix_cuda_step_host(picnix_arena_t* arena, double delt, int32_t nstep, double* uf,
                          double* uj, double* ff, double* xu, const int32_t* np_in,
                          const int32_t* np_cap, int32_t* np_out);

/* The two halves of picnix_cuda_step_host on their own: whole-rank state transfer through the same
 * copy/transposition pipeline.  They are the device side of the reference's snapshot and diagnostic
 * paths (PicChunk::pack/unpack, pic/pic_chunk.cpp:59-95; nix/statehandler.hpp; pic/diag/field.hpp,
 * particle.hpp), which read and write the host arrays -- and they work on multi-rank arenas, where
 * the caller drives the phases between them.  upload_state leaves the particles cell-ordered. */
int picnix_cuda_upload_state(picnix_arena_t* arena, double* uf, double* uj, double* ff, double* xu,
                             const int32_t* np_in, const int32_t* np_cap);
int picnix_cuda_download_state(picnix_arena_t* arena, double* uf, double* uj, double* ff,
                               double* xu, const int32_t* np_cap, int32_t* np_out);

/* ---- chunk moves between ranks (Application::rebalance, nix/application.hpp:332) ------------- */
/*
 * PicChunk::get_size_byte / pack / unpack (pic/pic_chunk.cpp:25-95) with the packed chunk kept in
 * DEVICE memory: fields (uf, uj, Friedman history) and the particles of every species of one local
 * chunk as one contiguous buffer that the caller hands to NCCL / a peer copy.  The receiving arena
 * is created with the new rank boundary (picnix_assign_rebalance) and capacities >= the incoming
 * counts; after all chunks arrived the caller runs picnix_cuda_sort_particle once.
 */
int picnix_cuda_chunk_pack_size(picnix_arena_t* arena, int32_t ichunk, int64_t* bytes);
int picnix_cuda_chunk_pack(picnix_arena_t* arena, int32_t ichunk, void* dev_buf, int64_t bytes);
int picnix_cuda_chunk_unpack(picnix_arena_t* arena, int32_t ichunk, const void* dev_buf,
                             int64_t bytes);

/* page-locked host memory for the arrays handed to picnix_cuda_step_host / upload / download
 * (an xt::xtensor can adopt it through xt::adapt); PICNIX_ERR_NODEVICE without a CUDA device */
int picnix_cuda_host_alloc(void** ptr, int64_t bytes);
int picnix_cuda_host_free(void* ptr);

#ifdef __cplusplus
}
#endif

#endif /* PICNIX_B200_H */
