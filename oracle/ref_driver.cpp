// -*- C++ -*-
// oracle/_ref driver (TEST INFRASTRUCTURE ONLY -- never linked into the product).
//
// Compiles the UNMODIFIED reference sources where they lie under /root/reference
// (pic/pic_chunk.cpp, nix/chunk.cpp, nix/chunkmap.cpp, nix/sfc.cpp, nix/balancer.cpp) against the
// single-process MPI shim in oracle/mpi_shim/ and exposes a flat C API so that tests and
// bench.py's reference arm can drive the reference's own PicChunk entry points
// (pic/pic_chunk.hpp:90-143) on caller-provided inputs.
//
// The only thing restated here is the per-step schedule of PicApplication::push_openmp
// (pic/pic_application.cpp:219-292), because PicApplication itself needs config files, MPI-IO
// diagnostics and a logger that are out of scope; the chunk-level calls, halo pack/unpack,
// MPI message flow, particle migration and sort are all the reference's own code.
#include "nix/balancer.hpp"
#include "nix/chunkmap.hpp"
#include "nix/random.hpp"
#include "pic_application.hpp"
#include "pic_chunk.hpp"
#include "pic_diag.hpp"

// The problem code of the reference's examples with physical boundaries, read in place and UNCHANGED:
// MainChunk::{setup, set_boundary_field, set_boundary_particle, inject_particle} of
// example/mrx/main.cpp (conducting walls in y, Harris sheet) and example/shock/main.cpp (wall at the
// lower x boundary, injection at the upper one).  Their headers are included above, so inside the
// namespaces only the example's own classes (and its main()) are declared.
namespace mrx_ex
{
#include "example/mrx/main.cpp"
}
namespace shock_ex
{
#include "example/shock/main.cpp"
}

#include <omp.h>

#include <cstdint>
#include <functional>
#include <memory>
#include <vector>

extern "C" void picnix_mpi_shim_reset(void);

namespace
{

struct RefConfig {
  int32_t ndims[3];    // global number of cells (z, y, x)
  int32_t cdims[3];    // number of chunks (z, y, x)
  int32_t periodic[3]; // periodicity (z, y, x)
  int32_t order;       // shape order 1..4
  int32_t pusher;      // 0 Boris, 1 Vay, 2 HigueraCary
  int32_t interp;      // 0 MC, 1 WT
  int32_t Ns;          // number of species
  int32_t vector_mode; // 0: reference 'scalar' kernels, 1: reference 'vector' kernels
  int32_t nthread;     // OpenMP threads for the chunk loops (<=0: all)
  int32_t problem;     // 0: plain PicChunk, 1: example/mrx MainChunk, 2: example/shock MainChunk
  double  cc;
  double  delh;
  double  friedman;
  double  buffer_ratio;
};

// What the driver needs from a chunk, whatever class it is built on: the PicChunk entry points and the
// (protected) arrays, reached through accessors the concrete class below provides.
struct RefArrays {
  xt::xtensor<float64, 4>* uf;
  xt::xtensor<float64, 4>* uj;
  xt::xtensor<float64, 5>* um;
  xt::xtensor<float64, 5>* ff;
  ParticleVec*             up;
  int*                     Ns;
  json*                    option;
};

class RefChunk
{
public:
  std::unique_ptr<PicChunk> chunk;
  std::function<RefArrays()> arrays;

  PicChunk& pic() { return *chunk; }
  auto&     ref_uf() { return *arrays().uf; }
  auto&     ref_uj() { return *arrays().uj; }
  auto&     ref_um() { return *arrays().um; }
  auto&     ref_ff() { return *arrays().ff; }
  auto&     ref_up() { return *arrays().up; }
  int       ref_Ns() { return *arrays().Ns; }
  json&     ref_option() { return *arrays().option; }

  // forwarding of the PicChunk entry points the driver calls
  void push_bfd(double dt) { pic().push_bfd(dt); }
  void push_efd(double dt) { pic().push_efd(dt); }
  void push_velocity(double dt) { pic().push_velocity(dt); }
  void push_position(double dt) { pic().push_position(dt); }
  void deposit_current(double dt) { pic().deposit_current(dt); }
  void deposit_moment() { pic().deposit_moment(); }
  void sort_particle(ParticleVec& p) { pic().sort_particle(p); }
  void init_friedman() { pic().init_friedman(); }
  void reset_load() { pic().reset_load(); }
  void set_boundary_pack(int mode) { pic().set_boundary_pack(mode); }
  void set_boundary_begin(int mode) { pic().set_boundary_begin(mode); }
  bool set_boundary_probe(int mode, bool wait) { return pic().set_boundary_probe(mode, wait); }
  void set_boundary_end(int mode) { pic().set_boundary_end(mode); }
  void set_boundary_unpack(int mode) { pic().set_boundary_unpack(mode); }
  void get_diverror(double& e, double& b) { pic().get_diverror(e, b); }
  void get_energy(double& e, double& b, double* p) { pic().get_energy(e, b, p); }
  int  get_nb_id(int dz, int dy, int dx) { return pic().get_nb_id(dz, dy, dx); }
  int  get_nb_rank(int dz, int dy, int dx) { return pic().get_nb_rank(dz, dy, dx); }
  int  get_boundary_margin() { return pic().get_boundary_margin(); }
};

// Base = PicChunk (periodic problems: the state comes through the C API) or an example's MainChunk
// (its boundary hooks apply; its own setup() runs when the configuration carries the example's
// parameters, otherwise the plain set-up below, so that any state can be fed to the example's hooks).
template <typename Base, bool IsExample>
class RefChunkT : public Base
{
public:
  using Base::Base;

  RefArrays arrays() { return {&this->uf, &this->uj, &this->um, &this->ff, &this->up, &this->Ns, &this->option}; }

  virtual void setup(json& config) override
  {
    if (IsExample && config.value("example_setup", false)) {
      Base::setup(config); // MainChunk::setup of the example: fields, particles, sort
      return;
    }
    PicChunk::setup(config); // pic/pic_chunk.cpp:135-262

    this->Ns     = config["Ns"].template get<int>();
    this->cc     = config["cc"].template get<float64>();
    float64 delh = config["delh"].template get<float64>();
    this->set_coordinate(delh, delh, delh);
    this->allocate();

    // same buffer setup as every example's MainChunk::setup (example/thermal/main.cpp:56-59)
    this->set_mpi_buffer(this->mpibufvec[BoundaryEmf], 0, 0, sizeof(float64) * 6);
    this->set_mpi_buffer(this->mpibufvec[BoundaryCur], 0, 0, sizeof(float64) * 4);
    this->set_mpi_buffer(this->mpibufvec[BoundaryMom], 0, 0, sizeof(float64) * this->Ns * 14);

    this->up.resize(this->Ns);
    for (int is = 0; is < this->Ns; is++) {
      this->up[is]     = std::make_shared<ParticleType>(0, *this);
      this->up[is]->Np = 0;
      this->up[is]->q  = 0;
      this->up[is]->m  = 1;
    }
    if (IsExample && config.contains("boundary"))
      this->option["boundary"] = config["boundary"]; // example/shock reads its wall / inflow values here
  }

  // plain chunks on non-periodic faces: the base class only logs an error (pic/pic_chunk.cpp:455-481)
  virtual void set_boundary_field(int mode) override
  {
    if (IsExample)
      Base::set_boundary_field(mode);
  }

};

template <typename T>
std::unique_ptr<RefChunk> make_chunk(const int dims[3], const bool has_dim[3], int id)
{
  auto typed   = std::make_unique<T>(dims, has_dim, id);
  T*   raw     = typed.get();
  auto wrapper = std::make_unique<RefChunk>();
  wrapper->chunk  = std::move(typed);
  wrapper->arrays = [raw]() { return raw->arrays(); };
  return wrapper;
}

using PlainChunk = RefChunkT<PicChunk, false>;
using MrxChunk   = RefChunkT<mrx_ex::MainChunk, true>;
using ShockChunk = RefChunkT<shock_ex::MainChunk, true>;

std::string g_problem_json; // parameters of the example's own setup(), set by ref_set_problem_json

struct RefSim {
  RefConfig                              cfg;
  std::unique_ptr<nix::ChunkMap>         chunkmap;
  std::vector<std::unique_ptr<RefChunk>> chunks;
  MPI_Comm                               comm[NumBoundaryMode][3][3][3];
  int                                    nthread;
};

template <typename F>
void for_each_chunk(RefSim* sim, F func)
{
  const int n = (int)sim->chunks.size();
#pragma omp parallel for schedule(dynamic) num_threads(sim->nthread)
  for (int i = 0; i < n; i++) {
    func(sim->chunks[i].get());
  }
}

void exchange(RefSim* sim, int mode)
{
  for_each_chunk(sim, [&](RefChunk* c) {
    c->set_boundary_pack(mode);
    c->set_boundary_begin(mode);
  });
  if (mode == BoundaryParticle) {
    for_each_chunk(sim, [&](RefChunk* c) { c->set_boundary_probe(mode, true); });
  }
  for_each_chunk(sim, [&](RefChunk* c) {
    c->set_boundary_end(mode);
    c->set_boundary_unpack(mode);
  });
}

} // namespace

extern "C" {

// JSON text with the example's own parameters; consumed by the next ref_create (problem 1 or 2), whose
// chunks then run the example's MainChunk::setup (key "example_setup": true) or only take its boundary
// values (key "boundary")
void ref_set_problem_json(const char* text)
{
  g_problem_json = text ? text : "";
}

void* ref_create(const RefConfig* cfg)
{
  static bool plog_ready = false;
  if (!plog_ready) {
    DebugPrinter::init();
    plog_ready = true;
  }

  auto sim     = new RefSim();
  sim->cfg     = *cfg;
  sim->nthread = cfg->nthread > 0 ? cfg->nthread : omp_get_max_threads();

  const int* nd = cfg->ndims;
  const int* cd = cfg->cdims;
  int        nc = cd[0] * cd[1] * cd[2];

  sim->chunkmap = std::make_unique<nix::ChunkMap>(cd[0], cd[1], cd[2]);
  sim->chunkmap->set_periodicity(cfg->periodic[0], cfg->periodic[1], cfg->periodic[2]);
  std::vector<int> boundary = {0, nc};
  sim->chunkmap->set_rank_boundary(boundary);

  // nix/application.cpp:262-271
  bool has_dim[3] = {
      (nd[0] == 1 && cd[0] == 1) ? false : true,
      (nd[1] == 1 && cd[1] == 1) ? false : true,
      (nd[2] == 1 && cd[2] == 1) ? false : true,
  };
  int dims[3] = {nd[0] / cd[0], nd[1] / cd[1], nd[2] / cd[2]};

  for (int mode = 0; mode < NumBoundaryMode; mode++)
    for (int iz = 0; iz < 3; iz++)
      for (int iy = 0; iy < 3; iy++)
        for (int ix = 0; ix < 3; ix++)
          MPI_Comm_dup(MPI_COMM_WORLD, &sim->comm[mode][iz][iy][ix]);

  static const char* pusher_name[3] = {"Boris", "Vay", "HigueraCary"};
  static const char* interp_name[2] = {"MC", "WT"};

  json config;
  config["Ns"]     = cfg->Ns;
  config["cc"]     = cfg->cc;
  config["delh"]   = cfg->delh;
  config["option"] = {{"vectorization", cfg->vector_mode ? "vector" : "scalar"},
                      {"order", cfg->order},
                      {"pusher", pusher_name[cfg->pusher]},
                      {"interpolation", interp_name[cfg->interp]},
                      {"seed_type", "fixed"},
                      {"friedman", cfg->friedman},
                      {"buffer_ratio", cfg->buffer_ratio}};

  if (!g_problem_json.empty()) {
    // the example's own parameters (config.toml [parameter]) on top of the driver's; its setup() runs
    json extra = json::parse(g_problem_json);
    for (auto it = extra.begin(); it != extra.end(); ++it) {
      if (it.key() == "option") {
        for (auto jt = it.value().begin(); jt != it.value().end(); ++jt)
          config["option"][jt.key()] = jt.value();
      } else {
        config[it.key()] = it.value();
      }
    }
    g_problem_json.clear();
  }

  for (int id = 0; id < nc; id++) {
    std::unique_ptr<RefChunk> chunk;
    if (cfg->problem == 1)
      chunk = make_chunk<MrxChunk>(dims, has_dim, id);
    else if (cfg->problem == 2)
      chunk = make_chunk<ShockChunk>(dims, has_dim, id);
    else
      chunk = make_chunk<PlainChunk>(dims, has_dim, id);

    // nix/chunkvector.hpp:56-81
    auto [cz, cy, cx] = sim->chunkmap->get_coordinate(id);
    for (int dirz = -1; dirz <= +1; dirz++) {
      for (int diry = -1; diry <= +1; diry++) {
        for (int dirx = -1; dirx <= +1; dirx++) {
          int nz   = sim->chunkmap->get_neighbor_coord(cz, dirz, 0);
          int ny   = sim->chunkmap->get_neighbor_coord(cy, diry, 1);
          int nx   = sim->chunkmap->get_neighbor_coord(cx, dirx, 2);
          int nbid = sim->chunkmap->get_chunkid(nz, ny, nx);
          chunk->pic().set_nb_id(dirz, diry, dirx, nbid);
          chunk->pic().set_nb_rank(dirz, diry, dirx, sim->chunkmap->get_rank(nbid));
        }
      }
    }

    // nix/application.cpp:291-300
    int offset[3] = {cz * nd[0] / cd[0], cy * nd[1] / cd[1], cx * nd[2] / cd[2]};
    chunk->pic().set_global_context(offset, nd);
    chunk->pic().setup(config);

    for (int mode = 0; mode < NumBoundaryMode; mode++)
      for (int iz = 0; iz < 3; iz++)
        for (int iy = 0; iy < 3; iy++)
          for (int ix = 0; ix < 3; ix++)
            chunk->pic().set_mpi_communicator(mode, iz, iy, ix, sim->comm[mode][iz][iy][ix]);

    sim->chunks.push_back(std::move(chunk));
  }

  return sim;
}

void ref_destroy(void* handle)
{
  delete static_cast<RefSim*>(handle);
}

int ref_num_chunks(void* handle)
{
  return (int)static_cast<RefSim*>(handle)->chunks.size();
}

int ref_num_threads(void* handle)
{
  return static_cast<RefSim*>(handle)->nthread;
}

// array shapes: shape[0..2] = padded (z, y, x) extents, shape[3] = boundary margin,
// shape[4] = Ng (size of pindex minus one)
void ref_get_shape(void* handle, int32_t* shape)
{
  auto  sim = static_cast<RefSim*>(handle);
  auto& uf  = sim->chunks[0]->ref_uf();
  shape[0]  = (int)uf.shape(0);
  shape[1]  = (int)uf.shape(1);
  shape[2]  = (int)uf.shape(2);
  shape[3]  = sim->chunks[0]->get_boundary_margin();
  shape[4]  = sim->chunks[0]->ref_up()[0]->Ng;
}

// chunkid[cz][cy][cx] and coord[id][3] (x, y, z) as stored by nix::ChunkMap
void ref_get_chunkmap(void* handle, int32_t* chunkid, int32_t* coord)
{
  auto       sim = static_cast<RefSim*>(handle);
  const int* cd  = sim->cfg.cdims;
  for (int cz = 0; cz < cd[0]; cz++)
    for (int cy = 0; cy < cd[1]; cy++)
      for (int cx = 0; cx < cd[2]; cx++)
        chunkid[(cz * cd[1] + cy) * cd[2] + cx] = sim->chunkmap->get_chunkid(cz, cy, cx);
  int nc = cd[0] * cd[1] * cd[2];
  for (int id = 0; id < nc; id++) {
    auto [cz, cy, cx] = sim->chunkmap->get_coordinate(id);
    coord[3 * id + 0] = cx;
    coord[3 * id + 1] = cy;
    coord[3 * id + 2] = cz;
  }
}

void ref_get_neighbors(void* handle, int ichunk, int32_t* nbid, int32_t* nbrank)
{
  auto sim = static_cast<RefSim*>(handle);
  for (int dz = -1; dz <= 1; dz++)
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++) {
        int k     = 9 * (dz + 1) + 3 * (dy + 1) + (dx + 1);
        nbid[k]   = sim->chunks[ichunk]->get_nb_id(dz, dy, dx);
        nbrank[k] = sim->chunks[ichunk]->get_nb_rank(dz, dy, dx);
      }
}

int ref_chunkmap_validate(void* handle)
{
  return static_cast<RefSim*>(handle)->chunkmap->validate() ? 1 : 0;
}

void ref_get_species(void* handle, int is, double* q, double* m)
{
  auto sim = static_cast<RefSim*>(handle);
  *q       = sim->chunks[0]->ref_up()[is]->q;
  *m       = sim->chunks[0]->ref_up()[is]->m;
}

void ref_set_species(void* handle, int is, double q, double m)
{
  auto sim = static_cast<RefSim*>(handle);
  for (auto& c : sim->chunks) {
    c->ref_up()[is]->q = q;
    c->ref_up()[is]->m = m;
  }
}

// which: 0 uf[..][6], 1 uj[..][4], 2 ff[..][3][6], 3 um[..][Ns][14]
static double* field_ptr(RefChunk* c, int which, size_t& size)
{
  switch (which) {
  case 0:
    size = c->ref_uf().size();
    return c->ref_uf().data();
  case 1:
    size = c->ref_uj().size();
    return c->ref_uj().data();
  case 2:
    size = c->ref_ff().size();
    return c->ref_ff().data();
  case 3:
    size = c->ref_um().size();
    return c->ref_um().data();
  }
  size = 0;
  return nullptr;
}

void ref_set_field(void* handle, int ichunk, int which, const double* src)
{
  auto    sim = static_cast<RefSim*>(handle);
  size_t  size;
  double* dst = field_ptr(sim->chunks[ichunk].get(), which, size);
  std::memcpy(dst, src, size * sizeof(double));
}

void ref_get_field(void* handle, int ichunk, int which, double* dst)
{
  auto    sim = static_cast<RefSim*>(handle);
  size_t  size;
  double* src = field_ptr(sim->chunks[ichunk].get(), which, size);
  std::memcpy(dst, src, size * sizeof(double));
}

// xu: AoS [np][7]; capacity follows the examples: ParticleType(np_alloc, chunk)
void ref_set_particles(void* handle, int ichunk, int is, const double* xu, int np, int np_alloc)
{
  auto  sim   = static_cast<RefSim*>(handle);
  auto  chunk = sim->chunks[ichunk].get();
  auto& up    = chunk->ref_up();
  double q = up[is]->q, m = up[is]->m;
  up[is]     = std::make_shared<ParticleType>(np_alloc, chunk->pic());
  up[is]->q  = q;
  up[is]->m  = m;
  up[is]->Np = np;
  std::memcpy(up[is]->xu.data(), xu, sizeof(double) * 7 * np);
}

int ref_get_np(void* handle, int ichunk, int is)
{
  return static_cast<RefSim*>(handle)->chunks[ichunk]->ref_up()[is]->Np;
}

int ref_get_np_total(void* handle, int ichunk, int is)
{
  return static_cast<RefSim*>(handle)->chunks[ichunk]->ref_up()[is]->Np_total;
}

// which: 0 xu, 1 xv; copies the first n rows
void ref_get_particles(void* handle, int ichunk, int is, int which, int n, double* dst)
{
  auto& p   = static_cast<RefSim*>(handle)->chunks[ichunk]->ref_up()[is];
  auto& arr = which == 0 ? p->xu : p->xv;
  std::memcpy(dst, arr.data(), sizeof(double) * 7 * n);
}

void ref_get_pindex(void* handle, int ichunk, int is, int32_t* dst)
{
  auto& p = static_cast<RefSim*>(handle)->chunks[ichunk]->ref_up()[is];
  std::memcpy(dst, p->pindex.data(), sizeof(int32_t) * (p->Ng + 1));
}

void ref_get_gindex(void* handle, int ichunk, int is, int n, int32_t* dst)
{
  auto& p = static_cast<RefSim*>(handle)->chunks[ichunk]->ref_up()[is];
  std::memcpy(dst, p->gindex.data(), sizeof(int32_t) * n);
}

// same sequence as the tail of MainChunk::setup + PicApplication::setup_chunks
// (example/thermal/main.cpp:61-111, pic/pic_application.cpp:106-130)
void ref_finalize_setup(void* handle)
{
  auto sim = static_cast<RefSim*>(handle);
  for_each_chunk(sim, [&](RefChunk* c) {
    c->init_friedman();
    c->sort_particle(c->ref_up());
  });
  exchange(sim, BoundaryEmf);
}

void ref_init_friedman(void* handle)
{
  for_each_chunk(static_cast<RefSim*>(handle), [&](RefChunk* c) { c->init_friedman(); });
}

void ref_push_bfd(void* handle, double delt)
{
  for_each_chunk(static_cast<RefSim*>(handle), [&](RefChunk* c) { c->push_bfd(delt); });
}

void ref_push_efd(void* handle, double delt)
{
  for_each_chunk(static_cast<RefSim*>(handle), [&](RefChunk* c) { c->push_efd(delt); });
}

void ref_push_velocity(void* handle, double delt)
{
  for_each_chunk(static_cast<RefSim*>(handle), [&](RefChunk* c) { c->push_velocity(delt); });
}

void ref_push_position(void* handle, double delt)
{
  for_each_chunk(static_cast<RefSim*>(handle), [&](RefChunk* c) { c->push_position(delt); });
}

void ref_deposit_current(void* handle, double delt)
{
  for_each_chunk(static_cast<RefSim*>(handle), [&](RefChunk* c) { c->deposit_current(delt); });
}

void ref_deposit_moment(void* handle)
{
  for_each_chunk(static_cast<RefSim*>(handle), [&](RefChunk* c) { c->deposit_moment(); });
}

void ref_sort_particle(void* handle)
{
  for_each_chunk(static_cast<RefSim*>(handle),
                 [&](RefChunk* c) { c->sort_particle(c->ref_up()); });
}

void ref_exchange(void* handle, int mode)
{
  exchange(static_cast<RefSim*>(handle), mode);
}

void ref_get_diverror(void* handle, int ichunk, double* efd, double* bfd)
{
  static_cast<RefSim*>(handle)->chunks[ichunk]->get_diverror(*efd, *bfd);
}

void ref_get_energy(void* handle, int ichunk, double* efd, double* bfd, double* particle)
{
  static_cast<RefSim*>(handle)->chunks[ichunk]->get_energy(*efd, *bfd, particle);
}

// PicApplication::push_openmp (pic/pic_application.cpp:219-292), repeated nstep times
void ref_step(void* handle, double delt, int nstep)
{
  auto      sim = static_cast<RefSim*>(handle);
  const int n   = (int)sim->chunks.size();

  for (int step = 0; step < nstep; step++) {
#pragma omp parallel num_threads(sim->nthread)
    {
#pragma omp for schedule(dynamic)
      for (int i = 0; i < n; i++) {
        auto chunk = sim->chunks[i].get();
        chunk->reset_load();
        chunk->push_bfd(0.5 * delt);
        chunk->push_velocity(delt);
        chunk->push_position(delt);
        chunk->deposit_current(delt);
        chunk->set_boundary_pack(BoundaryCur);
        chunk->set_boundary_begin(BoundaryCur);
        chunk->set_boundary_pack(BoundaryParticle);
        chunk->set_boundary_begin(BoundaryParticle);
        chunk->push_bfd(0.5 * delt);
      }

#pragma omp for schedule(dynamic)
      for (int i = 0; i < n; i++) {
        auto chunk = sim->chunks[i].get();
        chunk->set_boundary_end(BoundaryCur);
        chunk->set_boundary_unpack(BoundaryCur);
        chunk->push_efd(delt);
        chunk->set_boundary_pack(BoundaryEmf);
        chunk->set_boundary_begin(BoundaryEmf);
      }

#pragma omp for schedule(dynamic)
      for (int i = 0; i < n; i++) {
        sim->chunks[i]->set_boundary_probe(BoundaryParticle, true);
      }

#pragma omp for schedule(dynamic)
      for (int i = 0; i < n; i++) {
        auto chunk = sim->chunks[i].get();
        chunk->set_boundary_end(BoundaryParticle);
        chunk->set_boundary_unpack(BoundaryParticle);
      }

#pragma omp for schedule(dynamic)
      for (int i = 0; i < n; i++) {
        auto chunk = sim->chunks[i].get();
        chunk->set_boundary_end(BoundaryEmf);
        chunk->set_boundary_unpack(BoundaryEmf);
      }
    }
  }
}

} // extern "C"
