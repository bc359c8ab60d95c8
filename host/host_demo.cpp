// host_demo -- the C++ host mirror driving the device arena through the reference's loop nest.
//
//   host_demo [nstep]
//
// Builds a small 3-D thermal plasma (16^3 cells, 2x2x2 chunks, 2 species x 8 ppc), runs `nstep`
// steps twice on two arenas with identical initial state:
//   (a) picnix::host::push_openmp over PicChunkView objects from OpenMP workers
//       (first-caller-launches logic, the way a CudaPicChunk subclass of the reference would run)
//   (b) picnix_cuda_step (the whole-step entry point)
// and checks that both give the same particle counts and the same fields (to summation-order
// round-off of the current deposit).  Without a CUDA device the arena constructor throws
// PICNIX_ERR_NODEVICE: exit code 3, nothing is computed on the CPU.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "picnix_host.hpp"

using picnix::host::Arena;
using picnix::host::PicChunkView;

static void fill(Arena& A, const picnix_config_t& cfg, int ppc)
{
  const int Ns = cfg.Ns;
  const int dz = cfg.ndims[0] / cfg.cdims[0], dy = cfg.ndims[1] / cfg.cdims[1], dx = cfg.ndims[2] / cfg.cdims[2];
  const int Mz = A.padded[0], My = A.padded[1], Mx = A.padded[2];
  std::vector<int32_t> chunkid((size_t)cfg.cdims[0] * cfg.cdims[1] * cfg.cdims[2]), coord(chunkid.size() * 3);
  picnix_sfc_build(cfg.cdims[0], cfg.cdims[1], cfg.cdims[2], chunkid.data(), coord.data());

  const int npc = ppc * dz * dy * dx;
  std::vector<int32_t> cap((size_t)A.nchunk * Ns, (int32_t)(npc * 1.2));
  A.set_particle_capacity(cap);
  A.set_species(0, -1.0 / ppc, 1.0 / ppc);
  A.set_species(1, +1.0 / ppc, 25.0 / ppc);

  std::vector<double> uf((size_t)Mz * My * Mx * 6), xu((size_t)npc * 7);
  for (int ic = 0; ic < A.nchunk; ic++) {
    const int id = A.chunk_id_begin + ic;
    const double x0 = coord[id * 3 + 0] * dx * cfg.delx, y0 = coord[id * 3 + 1] * dy * cfg.dely,
                 z0 = coord[id * 3 + 2] * dz * cfg.delz;
    for (size_t i = 0; i < uf.size(); i += 6) {
      uf[i + 0] = uf[i + 1] = uf[i + 2] = 0.0;
      uf[i + 3] = 5.0; // uniform Bx
      uf[i + 4] = uf[i + 5] = 0.0;
    }
    A.upload_field(ic, PICNIX_FIELD_UF, uf.data());
    for (int is = 0; is < Ns; is++) {
      std::mt19937_64 rng(1000 * id + is);
      std::uniform_real_distribution<double> uni(0.0, 1.0);
      std::normal_distribution<double>       nrm(0.0, is == 0 ? 1.0 : 0.2);
      for (int ip = 0; ip < npc; ip++) {
        double* p = &xu[(size_t)ip * 7];
        p[0] = x0 + uni(rng) * dx * cfg.delx;
        p[1] = y0 + uni(rng) * dy * cfg.dely;
        p[2] = z0 + uni(rng) * dz * cfg.delz;
        p[3] = nrm(rng); p[4] = nrm(rng); p[5] = nrm(rng);
        int64_t pid = ((int64_t)id << 32) | (int64_t)(is * npc + ip);
        std::memcpy(&p[6], &pid, 8);
      }
      A.upload_particles(ic, is, xu.data(), npc);
    }
  }
  A.check(picnix_cuda_init_friedman(A.handle(), 0, -1));
  A.check(picnix_cuda_sort_particle(A.handle(), 0, -1));
  A.check(picnix_cuda_boundary_begin(A.handle(), PICNIX_BOUNDARY_EMF));
  A.check(picnix_cuda_boundary_end(A.handle(), PICNIX_BOUNDARY_EMF));
}

int main(int argc, char** argv)
{
  const int nstep = argc > 1 ? std::atoi(argv[1]) : 5;
  picnix_config_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  for (int i = 0; i < 3; i++) {
    cfg.ndims[i] = 16; cfg.cdims[i] = 2; cfg.periodic[i] = 1;
  }
  cfg.order = 2; cfg.pusher = PICNIX_PUSHER_BORIS; cfg.interp = PICNIX_INTERP_MC;
  cfg.Ns = 2; cfg.nrank = 1; cfg.rank = 0;
  cfg.cc = 10.0; cfg.delx = cfg.dely = cfg.delz = 1.0; cfg.friedman = 0.0; cfg.buffer_ratio = 0.2;
  const double delt = 0.05;

  try {
    Arena A(cfg), B(cfg);
    fill(A, cfg, 8);
    fill(B, cfg, 8);

    std::vector<PicChunkView> chunkvec;
    for (int ic = 0; ic < A.nchunk; ic++)
      chunkvec.emplace_back(A, ic);
    for (int s = 0; s < nstep; s++)
      picnix::host::push_openmp(chunkvec, delt);
    A.synchronize();

    B.check(picnix_cuda_step(B.handle(), delt, nstep));
    B.synchronize();

    auto npA = A.get_np(cfg.Ns), npB = B.get_np(cfg.Ns);
    long long totA = 0, totB = 0;
    bool same_np = true;
    for (size_t i = 0; i < npA.size(); i++) {
      totA += npA[i]; totB += npB[i];
      same_np = same_np && npA[i] == npB[i];
    }
    const size_t n = (size_t)A.padded[0] * A.padded[1] * A.padded[2] * 6;
    std::vector<double> fa(n), fb(n);
    double worst = 0.0, scale = 0.0;
    for (int ic = 0; ic < A.nchunk; ic++) {
      A.download_field(ic, PICNIX_FIELD_UF, fa.data());
      B.download_field(ic, PICNIX_FIELD_UF, fb.data());
      for (size_t i = 0; i < n; i++) {
        worst = std::fmax(worst, std::fabs(fa[i] - fb[i]));
        scale = std::fmax(scale, std::fabs(fb[i]));
      }
    }
    int64_t launches = 0, pushes = 0;
    picnix_cuda_get_counters(A.handle(), &launches, &pushes);
    const bool ok = same_np && totA == totB && worst <= 1e-11 * scale && launches > 0;
    std::printf("{\"nstep\": %d, \"particles\": %lld, \"same_np\": %s, \"max_field_diff\": %.3e, "
                "\"field_scale\": %.3e, \"kernel_launches\": %lld, \"ok\": %s}\n",
                nstep, totA, same_np ? "true" : "false", worst, scale, (long long)launches, ok ? "true" : "false");
    return ok ? 0 : 1;
  } catch (const picnix::host::Error& e) {
    std::fprintf(stderr, "picnix error %d: %s\n", e.status, e.what());
    return e.status == PICNIX_ERR_NODEVICE ? 3 : 2;
  }
}
