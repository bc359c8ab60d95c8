// The reference's example/mrx/main.cpp (Harris-sheet reconnection between conducting walls) -- MainChunk::setup,
// MainInterface, MainApplication (non-periodic y, load model), main(), all UNCHANGED and read in place from
// the reference tree -- with the chunk base class swapped for a CudaPicChunk that declares the problem's
// walls as boundary kinds of the arena.
//
// The example implements its walls as per-chunk host hooks (set_boundary_field / set_boundary_particle
// overrides, example/mrx/main.cpp:183-382).  On the device path those hooks must not run on the (stale)
// host mirrors during the step, so the two names are renamed inside main.cpp: the example's functions are
// still compiled and still used by MainChunk::setup() to initialise the host arrays' margins, but they no
// longer override the virtuals the application loop reaches; the step uses PICNIX_BC_CONDUCTING instead
// (csrc/boundary.cu, apply_particle_bc), which tests/test_gpu_boundaries.py pins against these very hooks.
// The headers main.cpp includes are included first, so that the macros rename nothing but main.cpp itself.
#include "cuda_pic_chunk.hpp"

#include "nix/random.hpp"
#include "pic_application.hpp"
#include "pic_chunk.hpp"
#include "pic_diag.hpp"

class MrxCudaChunk : public CudaPicChunk
{
public:
  using CudaPicChunk::CudaPicChunk;

  // what the example's hooks override after the renaming (never called by the application loop)
  virtual void mrx_host_set_boundary_field(int)
  {
  }
  virtual void mrx_host_set_boundary_particle(ParticleVec&)
  {
  }

  // conducting walls normal to y (example/mrx/main.cpp:183-382)
  virtual void declare_boundary_conditions(picnix_arena_t* arena) override
  {
    for (int side = 0; side < 2; side++) {
      if (picnix_cuda_set_boundary_condition(arena, 1, side, PICNIX_BC_CONDUCTING, nullptr) != PICNIX_OK) {
        ERROR << "picnix_b200: " << picnix_cuda_last_error(arena);
        MPI_Abort(MPI_COMM_WORLD, -1);
      }
    }
  }
};

#define PicChunk MrxCudaChunk
#define set_boundary_field mrx_host_set_boundary_field
#define set_boundary_particle mrx_host_set_boundary_particle
#include "example/mrx/main.cpp"
