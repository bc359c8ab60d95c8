// -*- C++ -*-
// picnix_host.hpp -- C++ host mirror of the reference's chunk interface on top of the C ABI.
//
// Header only, no reference headers, no CUDA headers: it is compiled with plain g++ against
// include/picnix_b200.h and linked with libpicnix_b200.so.  It shows (and the tests exercise) how the
// per-chunk virtuals of the reference map onto the batched device arena:
//
//   picnix::host::Arena         RAII over picnix_arena_t; every failing call throws Error with the
//                               message the reference would have logged before MPI_Abort
//                               (pic/pic_chunk.cpp:14-17)
//   picnix::host::PicChunkView  one per local chunk, with the method names of PicChunk
//                               (pic/pic_chunk.hpp:108-142).  The reference calls them once per
//                               chunk from OpenMP workers (pic/pic_application.cpp:219-292); here
//                               the first caller of a phase enqueues the arena-wide launch and the
//                               other chunks find it done.
//   picnix::host::push_openmp   the loop nest of PicApplication::push_openmp, verbatim in
//                               structure, over PicChunkView objects
#ifndef PICNIX_HOST_HPP
#define PICNIX_HOST_HPP

#include <cstdint>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/picnix_b200.h"

namespace picnix
{
namespace host
{

struct Error : std::runtime_error {
  int status;
  Error(int st, const std::string& msg) : std::runtime_error(msg), status(st) {}
};

// phases of one time step, in stream order
enum Phase {
  PhaseBfd1 = 0, PhasePush, PhaseCurBegin, PhaseParticleBegin, PhaseBfd2, PhaseCurEnd, PhaseEfd,
  PhaseEmfBegin, PhaseParticleEnd, PhaseEmfEnd, PhaseSort, NumPhase
};

class Arena
{
public:
  Arena(const picnix_config_t& cfg, const int32_t* boundary = nullptr)
  {
    int st = picnix_cuda_arena_create(&cfg, boundary, &h_);
    if (st != PICNIX_OK) {
      std::string msg = h_ ? picnix_cuda_last_error(h_) : "picnix_cuda_arena_create failed";
      if (st == PICNIX_ERR_NODEVICE && msg.empty())
        msg = "no CUDA device (there is no CPU fallback)";
      if (h_)
        picnix_cuda_arena_destroy(h_);
      h_ = nullptr;
      throw Error(st, msg);
    }
    check(picnix_cuda_get_layout(h_, &nchunk, &chunk_id_begin, padded, &margin, &Ng));
    for (int p = 0; p < NumPhase; p++)
      launched_[p] = -1;
  }
  ~Arena()
  {
    if (h_)
      picnix_cuda_arena_destroy(h_);
  }
  Arena(const Arena&)            = delete;
  Arena& operator=(const Arena&) = delete;

  picnix_arena_t* handle() { return h_; }
  void check(int st)
  {
    if (st != PICNIX_OK)
      throw Error(st, picnix_cuda_last_error(h_));
  }

  // true for exactly one caller per (step, phase); the launch happens under the lock, so a chunk
  // that finds the phase taken also finds it enqueued -- stream order == phase order
  template <class F>
  void once(int phase, F&& launch)
  {
    std::lock_guard<std::mutex> guard(mutex_);
    if (launched_[phase] == step)
      return;
    launched_[phase] = step;
    launch();
  }
  void next_step() { step++; }

  // state transfer ------------------------------------------------------------------------------
  void set_species(int is, double q, double m) { check(picnix_cuda_set_species(h_, is, q, m)); }
  void set_particle_capacity(const std::vector<int32_t>& cap)
  {
    check(picnix_cuda_set_particle_capacity(h_, cap.data()));
  }
  void upload_field(int ic, int which, const double* p) { check(picnix_cuda_upload_field(h_, ic, which, p)); }
  void download_field(int ic, int which, double* p) { check(picnix_cuda_download_field(h_, ic, which, p)); }
  void upload_particles(int ic, int is, const double* aos, int np)
  {
    check(picnix_cuda_upload_particles(h_, ic, is, aos, np));
  }
  void download_particles(int ic, int is, int n, double* aos)
  {
    check(picnix_cuda_download_particles(h_, ic, is, 0, n, aos));
  }
  std::vector<int32_t> get_np(int Ns)
  {
    std::vector<int32_t> np((size_t)nchunk * Ns);
    check(picnix_cuda_get_np(h_, np.data()));
    return np;
  }
  void synchronize() { check(picnix_cuda_synchronize(h_)); }

  int32_t nchunk = 0, chunk_id_begin = 0, padded[3] = {0, 0, 0}, margin = 0, Ng = 0;
  int64_t step = 0;

private:
  picnix_arena_t* h_ = nullptr;
  std::mutex      mutex_;
  int64_t         launched_[NumPhase];
};

// The chunk-level interface of the reference, forwarding to the arena.  Method names and the
// meaning of the arguments are PicChunk's (pic/pic_chunk.hpp:108-142, nix/chunk.hpp:386-543).
class PicChunkView
{
public:
  PicChunkView(Arena& arena, int local_index) : A(arena), local(local_index) {}

  void reset_load() { bfd_calls = 0; }
  void push_bfd(double delt)
  {
    const int phase = (bfd_calls++ % 2 == 0) ? PhaseBfd1 : PhaseBfd2;
    A.once(phase, [&] { A.check(picnix_cuda_push_bfd(A.handle(), 0, -1, delt)); });
  }
  // push_velocity + push_position + deposit_current are one fused launch (picnix_cuda_push_deposit_fused)
  void push_velocity(double delt)
  {
    A.once(PhasePush, [&] { A.check(picnix_cuda_push_deposit_fused(A.handle(), 0, -1, delt)); });
  }
  void push_position(double) {}
  void deposit_current(double) {}
  void push_efd(double delt)
  {
    A.once(PhaseEfd, [&] { A.check(picnix_cuda_push_efd(A.handle(), 0, -1, delt)); });
  }
  void set_boundary_pack(int) {} // pack + begin are one arena-wide call
  void set_boundary_begin(int mode)
  {
    A.once(begin_phase(mode), [&] { A.check(picnix_cuda_boundary_begin(A.handle(), mode)); });
  }
  bool set_boundary_probe(int, bool) { return true; }
  void set_boundary_end(int mode)
  {
    A.once(end_phase(mode), [&] { A.check(picnix_cuda_boundary_end(A.handle(), mode)); });
  }
  void set_boundary_unpack(int) {}
  void sort_particle()
  {
    A.once(PhaseSort, [&] { A.check(picnix_cuda_sort_particle(A.handle(), 0, -1)); });
  }

  Arena& A;
  int    local;

private:
  int        bfd_calls = 0;
  static int begin_phase(int mode)
  {
    return mode == PICNIX_BOUNDARY_CUR ? PhaseCurBegin
                                       : (mode == PICNIX_BOUNDARY_PARTICLE ? PhaseParticleBegin : PhaseEmfBegin);
  }
  static int end_phase(int mode)
  {
    return mode == PICNIX_BOUNDARY_CUR ? PhaseCurEnd
                                       : (mode == PICNIX_BOUNDARY_PARTICLE ? PhaseParticleEnd : PhaseEmfEnd);
  }
};

// PicApplication::push_openmp (pic/pic_application.cpp:219-292): same five loops over the chunks,
// same calls in the same order; single-rank arenas (multi-rank callers move the peer buffers
// between begin and end)
inline void push_openmp(std::vector<PicChunkView>& chunkvec, double delt)
{
  const int n = (int)chunkvec.size();
#pragma omp parallel
  {
#pragma omp for schedule(dynamic)
    for (int i = 0; i < n; i++) {
      PicChunkView& chunk = chunkvec[i];
      chunk.reset_load();
      chunk.push_bfd(0.5 * delt);
      chunk.push_velocity(delt);
      chunk.push_position(delt);
      chunk.deposit_current(delt);
      chunk.set_boundary_pack(PICNIX_BOUNDARY_CUR);
      chunk.set_boundary_begin(PICNIX_BOUNDARY_CUR);
      chunk.set_boundary_pack(PICNIX_BOUNDARY_PARTICLE);
      chunk.set_boundary_begin(PICNIX_BOUNDARY_PARTICLE);
      chunk.push_bfd(0.5 * delt);
    }
#pragma omp for schedule(dynamic)
    for (int i = 0; i < n; i++) {
      PicChunkView& chunk = chunkvec[i];
      chunk.set_boundary_end(PICNIX_BOUNDARY_CUR);
      chunk.set_boundary_unpack(PICNIX_BOUNDARY_CUR);
      chunk.push_efd(delt);
      chunk.set_boundary_pack(PICNIX_BOUNDARY_EMF);
      chunk.set_boundary_begin(PICNIX_BOUNDARY_EMF);
    }
#pragma omp for schedule(dynamic)
    for (int i = 0; i < n; i++)
      chunkvec[i].set_boundary_probe(PICNIX_BOUNDARY_PARTICLE, true);
#pragma omp for schedule(dynamic)
    for (int i = 0; i < n; i++) {
      chunkvec[i].set_boundary_end(PICNIX_BOUNDARY_PARTICLE);
      chunkvec[i].set_boundary_unpack(PICNIX_BOUNDARY_PARTICLE);
    }
#pragma omp for schedule(dynamic)
    for (int i = 0; i < n; i++) {
      chunkvec[i].set_boundary_end(PICNIX_BOUNDARY_EMF);
      chunkvec[i].set_boundary_unpack(PICNIX_BOUNDARY_EMF);
    }
  }
  if (n > 0)
    chunkvec[0].A.next_step();
}

} // namespace host
} // namespace picnix

#endif
