// -*- C++ -*-
// CudaPicChunk: the reference's PicChunk (pic/pic_chunk.hpp:12-143) with its per-timestep kernels
// forwarded to the B200 library through the C ABI (include/picnix_b200.h).
//
// This header is compiled AGAINST THE REFERENCE'S OWN HEADERS (amanotk/pic-nix): it is the file a
// maintainer adds next to pic/pic_chunk.hpp.  A problem's MainChunk derives from CudaPicChunk instead
// of PicChunk (example/thermal/main.cpp:10-112 is used unchanged by host/ref_binding/Makefile, with
// the base class swapped on the compiler command line) and the application loop
// (nix::Application::main, nix/application.cpp:50-87; PicApplication::push_openmp,
// pic/pic_application.cpp:219-292) runs as it is.
//
// How per-chunk virtual calls become batched launches.  The reference calls every entry point once
// per chunk from OpenMP workers; the library wants ONE call per phase for all chunks of the rank
// (a launch per 16^3-cell chunk is hopeless on a GPU).  Every chunk counts how often it has called
// each entry point; the hub remembers how often the phase has been launched.  The first chunk whose
// count exceeds the hub's launches the phase for everybody, the others find it done:
//
//     chunk 0: push_bfd #1 -> launch          chunk 1: push_bfd #1 -> already launched, return
//     chunk 0: push_bfd #2 -> launch          ...
//
// Because every worker issues its calls in program order and launches are serialised by the hub's
// mutex, the stream sees the phases in the order of PicApplication::push_openmp whatever the
// interleaving of the workers.  push_velocity / push_position are deferred until deposit_current so that
// the three become the fused kernel (picnix_cuda_push_deposit_fused).
//
// Host mirrors.  setup() fills the host arrays as in the reference; they are uploaded when the first
// phase is called.  Diagnostics that read the host arrays through get_internal_data() (not virtual,
// pic/pic_chunk.hpp:49-75) see the state of the last host synchronisation: every
// `PICNIX_SYNC_HOST_INTERVAL` steps at the end of push (default 1 = every step; 0 = never) and before
// pack().  get_energy / get_diverror are served from the device and never need the mirrors.
//
// Scope: one rank (all chunks in one arena).  Chunks owned by other ranks need the per-peer buffers of
// picnix_cuda_get_comm_buffer moved by MPI between begin and end; see INTEGRATION.md.
#ifndef PICNIX_B200_CUDA_PIC_CHUNK_HPP
#define PICNIX_B200_CUDA_PIC_CHUNK_HPP

#include "pic_chunk.hpp" // the reference's

#include "picnix_b200.h"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

class CudaPicChunk;

namespace picnix_binding
{

enum PhaseKey {
  K_BFD = 0,
  K_EFD,
  K_VEL,
  K_POS,
  K_CUR,
  K_MOM,
  K_SORT,
  K_BEGIN,               // + boundary mode (4)
  K_END = K_BEGIN + 4,   // + boundary mode (4)
  K_COUNT = K_END + 4
};

/// One per process: owns the arena of the rank and turns per-chunk calls into per-phase launches.
class ArenaHub
{
public:
  std::mutex                 mtx;
  std::vector<CudaPicChunk*> chunks; // registered by CudaPicChunk::setup, ordered by chunk id when live
  picnix_arena_t*            arena = nullptr;
  int64_t                    gen[K_COUNT] = {};
  int64_t                    epoch        = 0; // bumped by every launch: invalidates the diagnostic cache
  int64_t                    step         = 0;
  int                        sync_interval = 1;
  bool                       pending_vel = false, pending_pos = false;
  double                     delt_vel = 0, delt_pos = 0;
  // diagnostics cache (per chunk): div E - rho, div B, E^2/2, B^2/2, particle energies
  int64_t             diag_epoch = -1, energy_epoch = -1;
  std::vector<double> dive, divb, ene_e, ene_b, ene_p;

  static ArenaHub& instance()
  {
    static ArenaHub hub;
    return hub;
  }

  void check(int status, const char* what)
  {
    if (status != PICNIX_OK) {
      ERROR << tfm::format("picnix_b200: %s failed: %s", what,
                           arena ? picnix_cuda_last_error(arena) : "no arena");
      MPI_Abort(MPI_COMM_WORLD, -1);
    }
  }

  inline void make_live();                     // create the arena from the registered chunks, upload
  inline void sync_host(CudaPicChunk* only);   // device -> host mirrors (all chunks, or one)
  inline void sync_moment_host();
  inline void refresh_diag(bool with_particle);

  // launch what push_velocity / push_position deferred
  void flush_pending()
  {
    if (pending_vel) {
      check(picnix_cuda_push_velocity(arena, 0, -1, delt_vel), "push_velocity");
      pending_vel = false;
    }
    if (pending_pos) {
      check(picnix_cuda_push_position(arena, 0, -1, delt_pos), "push_position");
      pending_pos = false;
    }
  }

  ~ArenaHub()
  {
    if (arena != nullptr)
      picnix_cuda_arena_destroy(arena);
  }
};

} // namespace picnix_binding

class CudaPicChunk : public PicChunk
{
  friend class picnix_binding::ArenaHub;

protected:
  int64_t calls[picnix_binding::K_COUNT] = {};
  bool    registered                     = false;

  using Hub = picnix_binding::ArenaHub;

  /// the n-th call of entry point `key` by this chunk launches the phase iff nobody has yet
  template <typename F>
  void phase(int key, bool keep_pending, F&& launch)
  {
    Hub&                        hub = Hub::instance();
    std::lock_guard<std::mutex> lock(hub.mtx);
    if (hub.arena == nullptr)
      hub.make_live();
    if (++calls[key] > hub.gen[key]) {
      if (!keep_pending)
        hub.flush_pending();
      launch(hub);
      hub.gen[key] = calls[key];
      hub.epoch++;
    }
  }

public:
  CudaPicChunk(const int dims[3], const bool has_dim[3], int id = 0) : PicChunk(dims, has_dim, id)
  {
  }

  /// Physical boundaries of the problem.  An example whose MainChunk overrides set_boundary_field /
  /// set_boundary_particle (example/mrx, example/shock) declares them here once instead, as boundary
  /// kinds of the arena (picnix_cuda_set_boundary_condition); the device applies them after every halo
  /// exchange and inside every position push.  Called on one chunk right after the arena is created.
  virtual void declare_boundary_conditions(picnix_arena_t* arena)
  {
    (void)arena;
  }

  virtual ~CudaPicChunk() override
  {
    Hub&                        hub = Hub::instance();
    std::lock_guard<std::mutex> lock(hub.mtx);
    auto it = std::find(hub.chunks.begin(), hub.chunks.end(), this);
    if (it != hub.chunks.end())
      hub.chunks.erase(it);
    if (hub.chunks.empty() && hub.arena != nullptr) {
      picnix_cuda_arena_destroy(hub.arena);
      hub.arena = nullptr;
      std::fill(hub.gen, hub.gen + picnix_binding::K_COUNT, 0);
    }
  }

  /// PicChunk::setup parses `option` (pic/pic_chunk.cpp:135-262); the problem's setup() then fills the
  /// host arrays.  The chunk only registers itself here: the arena is created at the first phase.
  virtual void setup(json& config) override
  {
    PicChunk::setup(config);
    Hub&                        hub = Hub::instance();
    std::lock_guard<std::mutex> lock(hub.mtx);
    if (hub.arena != nullptr) {
      ERROR << "CudaPicChunk: chunks cannot be added to a live arena (rebalance across ranks is driven "
               "by picnix_cuda_chunk_pack/unpack, see INTEGRATION.md)";
      MPI_Abort(MPI_COMM_WORLD, -1);
    }
    if (!registered) {
      hub.chunks.push_back(this);
      registered = true;
    }
  }

  // ---- kernels (pic/pic_chunk.hpp:131-142) ----------------------------------------------------
  virtual void push_bfd(float64 delt) override
  {
    phase(picnix_binding::K_BFD, false,
          [&](Hub& h) { h.check(picnix_cuda_push_bfd(h.arena, 0, -1, delt), "push_bfd"); });
  }

  virtual void push_efd(float64 delt) override
  {
    phase(picnix_binding::K_EFD, false,
          [&](Hub& h) { h.check(picnix_cuda_push_efd(h.arena, 0, -1, delt), "push_efd"); });
  }

  virtual void push_velocity(float64 delt) override
  {
    phase(picnix_binding::K_VEL, true, [&](Hub& h) {
      h.flush_pending();
      h.pending_vel = true;
      h.delt_vel    = delt;
    });
  }

  virtual void push_position(float64 delt) override
  {
    phase(picnix_binding::K_POS, true, [&](Hub& h) {
      if (h.pending_pos)
        h.flush_pending();
      h.pending_pos = true;
      h.delt_pos    = delt;
    });
  }

  virtual void deposit_current(float64 delt) override
  {
    phase(picnix_binding::K_CUR, true, [&](Hub& h) {
      if (h.pending_vel && h.pending_pos && h.delt_vel == delt && h.delt_pos == delt) {
        // velocity + position + count + deposit in one pass over the particles
        h.pending_vel = h.pending_pos = false;
        h.check(picnix_cuda_push_deposit_fused(h.arena, 0, -1, delt), "push_deposit_fused");
      } else {
        h.flush_pending();
        h.check(picnix_cuda_deposit_current(h.arena, 0, -1, delt), "deposit_current");
      }
    });
  }

  virtual void deposit_moment() override
  {
    phase(picnix_binding::K_MOM, false,
          [&](Hub& h) { h.check(picnix_cuda_deposit_moment(h.arena), "deposit_moment"); });
  }

  virtual void sort_particle(ParticleVec& particle) override
  {
    Hub& hub = Hub::instance();
    {
      std::lock_guard<std::mutex> lock(hub.mtx);
      if (hub.arena == nullptr) {
        PicChunk::sort_particle(particle); // still in setup(): host arrays
        return;
      }
    }
    phase(picnix_binding::K_SORT, false,
          [&](Hub& h) { h.check(picnix_cuda_sort_particle(h.arena, 0, -1), "sort_particle"); });
  }

  // ---- boundary exchange (pic/pic_chunk.cpp:271-405): neighbours inside the arena need no messages ----
  virtual void set_boundary_pack(int mode) override
  {
    if (mode == BoundaryParticle)
      this->inject_particle(up); // host hook, as PicChunk::set_boundary_pack does
  }

  virtual void set_boundary_begin(int mode) override
  {
    phase(picnix_binding::K_BEGIN + mode, false,
          [&](Hub& h) { h.check(picnix_cuda_boundary_begin(h.arena, mode), "boundary_begin"); });
  }

  virtual bool set_boundary_probe(int mode, bool wait) override
  {
    return true;
  }

  virtual void set_boundary_end(int mode) override
  {
    phase(picnix_binding::K_END + mode, false, [&](Hub& h) {
      h.check(picnix_cuda_boundary_end(h.arena, mode), "boundary_end");
      if (mode == BoundaryEmf) {
        // the last device phase of PicApplication::push_openmp: refresh the host mirrors if due
        h.step++;
        if (h.sync_interval > 0 && h.step % h.sync_interval == 0)
          h.sync_host(nullptr);
      } else if (mode == BoundaryMom && h.sync_interval > 0) {
        h.sync_moment_host();
      }
    });
  }

  virtual void set_boundary_unpack(int mode) override
  {
    this->set_boundary_field(mode); // physical boundary hook, as PicChunk::set_boundary_unpack does
  }

  // ---- diagnostics (pic/pic_chunk.cpp:407-445) ---------------------------------------------------
  virtual void get_diverror(float64& efd, float64& bfd) override
  {
    Hub&                        hub = Hub::instance();
    std::lock_guard<std::mutex> lock(hub.mtx);
    if (hub.arena == nullptr)
      hub.make_live();
    hub.refresh_diag(false);
    efd = hub.dive[this->myid];
    bfd = hub.divb[this->myid];
  }

  virtual void get_energy(float64& efd, float64& bfd, float64 particle[]) override
  {
    Hub&                        hub = Hub::instance();
    std::lock_guard<std::mutex> lock(hub.mtx);
    if (hub.arena == nullptr)
      hub.make_live();
    hub.refresh_diag(true);
    efd = hub.ene_e[this->myid];
    bfd = hub.ene_b[this->myid];
    for (int is = 0; is < Ns; is++)
      particle[is] = hub.ene_p[this->myid * Ns + is];
  }

  // ---- snapshot / rebalance serialisation reads the host arrays: refresh them first ----------------
  virtual int pack(void* buffer, int address) override
  {
    Hub& hub = Hub::instance();
    {
      std::lock_guard<std::mutex> lock(hub.mtx);
      if (hub.arena != nullptr)
        hub.sync_host(this);
    }
    return PicChunk::pack(buffer, address);
  }
};

namespace picnix_binding
{

inline void ArenaHub::make_live()
{
  if (chunks.empty()) {
    ERROR << "CudaPicChunk: no chunk registered";
    MPI_Abort(MPI_COMM_WORLD, -1);
  }
  int nprocess = 1;
  MPI_Comm_size(MPI_COMM_WORLD, &nprocess);
  if (nprocess != 1) {
    ERROR << "CudaPicChunk binds one rank; multi-rank runs move the per-peer buffers of "
             "picnix_cuda_get_comm_buffer with MPI (INTEGRATION.md)";
    MPI_Abort(MPI_COMM_WORLD, -1);
  }
  if (const char* env = std::getenv("PICNIX_SYNC_HOST_INTERVAL"))
    sync_interval = std::atoi(env);

  std::sort(chunks.begin(), chunks.end(),
            [](const CudaPicChunk* a, const CudaPicChunk* b) { return a->myid < b->myid; });
  CudaPicChunk*   c0 = chunks[0];
  picnix_config_t cfg;
  for (int d = 0; d < 3; d++) {
    cfg.ndims[d]    = c0->gdims[d];
    cfg.cdims[d]    = c0->gdims[d] / c0->dims[d];
    cfg.periodic[d] = 1;
  }
  // a chunk on the lower face of a non-periodic direction has no neighbour there (MPI_PROC_NULL)
  for (CudaPicChunk* c : chunks) {
    if (c->offset[0] == 0 && c->get_nb_rank(-1, 0, 0) == MPI_PROC_NULL)
      cfg.periodic[0] = 0;
    if (c->offset[1] == 0 && c->get_nb_rank(0, -1, 0) == MPI_PROC_NULL)
      cfg.periodic[1] = 0;
    if (c->offset[2] == 0 && c->get_nb_rank(0, 0, -1) == MPI_PROC_NULL)
      cfg.periodic[2] = 0;
  }
  cfg.order        = c0->order;
  cfg.pusher       = c0->option["pusher"].get<int>();
  cfg.interp       = c0->option["interpolation"].get<int>();
  cfg.Ns           = c0->Ns;
  cfg.nrank        = 1;
  cfg.rank         = 0;
  cfg.cc           = c0->cc;
  cfg.delx         = c0->delx;
  cfg.dely         = c0->dely;
  cfg.delz         = c0->delz;
  cfg.friedman     = c0->option.value("friedman", 0.0);
  cfg.buffer_ratio = c0->option.value("buffer_ratio", 0.2);
  check(picnix_cuda_arena_create(&cfg, nullptr, &arena), "arena_create");
  c0->declare_boundary_conditions(arena);

  int32_t nchunk = 0, begin = 0;
  check(picnix_cuda_get_layout(arena, &nchunk, &begin, nullptr, nullptr, nullptr), "get_layout");
  if (nchunk != (int)chunks.size() || begin != 0) {
    ERROR << tfm::format("CudaPicChunk: arena owns %d chunks, the application registered %d", nchunk,
                         (int)chunks.size());
    MPI_Abort(MPI_COMM_WORLD, -1);
  }
  // the arena's space-filling-curve map must be the application's (bit-exact integer logic)
  for (int ic = 0; ic < nchunk; ic++) {
    int32_t nbid[27];
    check(picnix_cuda_get_neighbors(arena, ic, nbid, nullptr), "get_neighbors");
    CudaPicChunk* c = chunks[ic];
    for (int k = 0; k < 27; k++) {
      const int dz = k / 9 - 1, dy = (k / 3) % 3 - 1, dx = k % 3 - 1;
      const int ref = c->get_nb_rank(dz, dy, dx) == MPI_PROC_NULL ? -1 : c->get_nb_id(dz, dy, dx);
      if (c->myid != ic || (ref >= 0 && nbid[k] != ref)) {
        ERROR << tfm::format("CudaPicChunk: chunk map mismatch at chunk %d direction %d", ic, k);
        MPI_Abort(MPI_COMM_WORLD, -1);
      }
    }
  }

  const int            Ns = c0->Ns;
  std::vector<int32_t> cap(nchunk * Ns);
  for (int ic = 0; ic < nchunk; ic++)
    for (int is = 0; is < Ns; is++)
      cap[ic * Ns + is] = chunks[ic]->up[is]->Np_total;
  for (int is = 0; is < Ns; is++)
    check(picnix_cuda_set_species(arena, is, c0->up[is]->q, c0->up[is]->m), "set_species");
  check(picnix_cuda_set_particle_capacity(arena, cap.data()), "set_particle_capacity");
  for (int ic = 0; ic < nchunk; ic++) {
    CudaPicChunk* c = chunks[ic];
    check(picnix_cuda_upload_field(arena, ic, PICNIX_FIELD_UF, c->uf.data()), "upload uf");
    check(picnix_cuda_upload_field(arena, ic, PICNIX_FIELD_UJ, c->uj.data()), "upload uj");
    check(picnix_cuda_upload_field(arena, ic, PICNIX_FIELD_FF, c->ff.data()), "upload ff");
    for (int is = 0; is < Ns; is++)
      check(picnix_cuda_upload_particles(arena, ic, is, c->up[is]->xu.data(), c->up[is]->Np),
            "upload particles");
  }
  check(picnix_cuda_sort_particle(arena, 0, -1), "sort_particle"); // keys and pindex on the device
}

inline void ArenaHub::sync_host(CudaPicChunk* only)
{
  flush_pending();
  const int            Ns = chunks[0]->Ns;
  std::vector<int32_t> np(chunks.size() * Ns);
  check(picnix_cuda_get_np(arena, np.data()), "get_np");
  for (size_t ic = 0; ic < chunks.size(); ic++) {
    CudaPicChunk* c = chunks[ic];
    if (only != nullptr && c != only)
      continue;
    check(picnix_cuda_download_field(arena, (int)ic, PICNIX_FIELD_UF, c->uf.data()), "download uf");
    check(picnix_cuda_download_field(arena, (int)ic, PICNIX_FIELD_UJ, c->uj.data()), "download uj");
    check(picnix_cuda_download_field(arena, (int)ic, PICNIX_FIELD_FF, c->ff.data()), "download ff");
    for (int is = 0; is < Ns; is++) {
      const int n = np[ic * Ns + is];
      if (n > c->up[is]->Np_total)
        c->up[is]->resize(n); // XtensorParticle::resize, nix/xtensor_particle.hpp:70-115
      c->up[is]->Np = n;
      check(picnix_cuda_download_particles(arena, (int)ic, is, 0, n, c->up[is]->xu.data()),
            "download particles");
    }
  }
}

inline void ArenaHub::sync_moment_host()
{
  for (size_t ic = 0; ic < chunks.size(); ic++)
    check(picnix_cuda_download_field(arena, (int)ic, PICNIX_FIELD_UM, chunks[ic]->um.data()), "download um");
}

inline void ArenaHub::refresh_diag(bool with_particle)
{
  const size_t n  = chunks.size();
  const int    Ns = chunks[0]->Ns;
  if (diag_epoch != epoch) {
    flush_pending();
    dive.resize(n);
    divb.resize(n);
    ene_e.resize(n);
    ene_b.resize(n);
    check(picnix_cuda_get_diverror(arena, dive.data(), divb.data()), "get_diverror");
    check(picnix_cuda_get_field_energy(arena, ene_e.data(), ene_b.data()), "get_field_energy");
    diag_epoch = epoch;
  }
  if (with_particle && energy_epoch != epoch) {
    // needs deposit_moment + the BoundaryMom exchange, which PicApplication::calculate_moment ran
    ene_p.resize(n * Ns);
    check(picnix_cuda_get_particle_energy(arena, ene_p.data()), "get_particle_energy");
    energy_epoch = epoch;
  }
}

} // namespace picnix_binding

#endif
