// host_nccl_demo -- the thermal-plasma benchmark problem driven from C++ on N GPUs of one node:
// one process per GPU (launch: python -m torch.distributed.run --no-python --nproc-per-node N
// --master-addr 127.0.0.1 host/host_nccl_demo [cells per rank] [steps]), chunk ids split over the ranks
// like the reference's MPI ranks, halos and particle migration over NCCL (picnix_host_nccl.hpp).
//
// The initial condition is example/thermal/main.cpp:69-108: std::mt19937_64 seeded with the chunk id,
// positions uniform in the chunk (the same stream for both species: charge neutrality), velocities
// normal(0, vt).  Rank 0 prints one JSON line: particle-steps/s over all ranks (CUDA events on the compute
// stream, max over ranks), particle conservation, charge-conservation residual.
#include <cinttypes>
#include <cmath>
#include <cstring>
#include <random>

#include "picnix_host_nccl.hpp"

using namespace picnix::host;

int main(int argc, char** argv)
{
  const int cells = argc > 1 ? std::atoi(argv[1]) : 128;
  const int steps = argc > 2 ? std::atoi(argv[2]) : 20;
  const int ppc = 32, chunk = 16, Ns = 2;
  const double delt = 0.05, delh = 1.0, cc = 10.0, Bx = 5.0;
  const double qm[2] = {-1.0, +0.1}, ro[2] = {1.0, 10.0}, vt[2] = {1.0, 0.31622776601};

  RankEnv env;
  try {
    cuda_check(cudaSetDevice(env.local), "cudaSetDevice");
    // ranks arranged (1,1,1) (1,1,2) (1,2,2) (2,2,2): every rank's share is a cells^3 block of the box
    int lay[3] = {1, 1, 1};
    for (int n = env.world, d = 2; n > 1; n /= 2, d = (d + 2) % 3)
      lay[d] *= 2;
    picnix_config_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    for (int i = 0; i < 3; i++) {
      cfg.ndims[i]    = cells * lay[i];
      cfg.cdims[i]    = cfg.ndims[i] / chunk;
      cfg.periodic[i] = 1;
    }
    cfg.order = 2;
    cfg.pusher = PICNIX_PUSHER_BORIS;
    cfg.interp = PICNIX_INTERP_MC;
    cfg.Ns = Ns;
    cfg.nrank = env.world;
    cfg.rank = env.rank;
    cfg.cc = cc;
    cfg.delx = cfg.dely = cfg.delz = delh;
    cfg.buffer_ratio = 0.2;

    Arena A(cfg);
    const char* port = std::getenv("MASTER_PORT");
    NcclTransport T(A, env, std::string("/tmp/picnix_nccl_id_") + (port ? port : "0"));

    const int nglobal = cfg.cdims[0] * cfg.cdims[1] * cfg.cdims[2];
    std::vector<int32_t> chunkid(nglobal), coord(3 * (size_t)nglobal);
    A.check(picnix_sfc_build(cfg.cdims[0], cfg.cdims[1], cfg.cdims[2], chunkid.data(), coord.data()));

    const int mp = ppc * chunk * chunk * chunk;
    std::vector<int32_t> cap((size_t)A.nchunk * Ns, (int32_t)(mp * 1.2));
    for (int is = 0; is < Ns; is++)
      A.set_species(is, qm[is] * ro[is] / ppc, ro[is] / ppc);
    A.set_particle_capacity(cap);

    const size_t ncell = (size_t)A.padded[0] * A.padded[1] * A.padded[2];
    std::vector<double> uf(ncell * 6, 0.0), xu((size_t)mp * 7);
    for (size_t c = 0; c < ncell; c++)
      uf[c * 6 + 3] = Bx;
    for (int ic = 0; ic < A.nchunk; ic++) {
      const int gid = A.chunk_id_begin + ic;
      A.upload_field(ic, PICNIX_FIELD_UF, uf.data());
      const double lo[3] = {coord[3 * gid + 0] * chunk * delh, coord[3 * gid + 1] * chunk * delh,
                            coord[3 * gid + 2] * chunk * delh}; // x, y, z
      std::mt19937_64 mtv(gid);
      std::normal_distribution<double> normal(0.0, 1.0);
      for (int is = 0; is < Ns; is++) {
        std::mt19937_64 mtp(gid); // same positions for every species
        std::uniform_real_distribution<double> uniform(0.0, 1.0);
        for (int ip = 0; ip < mp; ip++) {
          double* p = &xu[(size_t)ip * 7];
          for (int k = 0; k < 3; k++)
            p[k] = uniform(mtp) * chunk * delh + lo[k];
          for (int k = 3; k < 6; k++)
            p[k] = normal(mtv) * vt[is];
          const int64_t id = (int64_t)mp * gid + ip;
          std::memcpy(&p[6], &id, sizeof(id));
        }
        A.upload_particles(ic, is, xu.data(), mp);
      }
    }
    A.check(picnix_cuda_init_friedman(A.handle(), 0, -1));
    A.check(picnix_cuda_sort_particle(A.handle(), 0, -1));
    exchange(A, T, PICNIX_BOUNDARY_EMF);

    auto total_np = [&]() {
      double n = 0;
      for (int32_t v : A.get_np(Ns))
        n += v;
      return T.allreduce(n, ncclSum);
    };
    const double np0 = total_np();
    for (int k = 0; k < 5; k++)
      step_phases(A, T, delt);
    cuda_check(cudaStreamSynchronize(T.compute), "sync");
    T.allreduce(0.0, ncclSum); // barrier
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, T.compute);
    for (int k = 0; k < steps; k++)
      step_phases(A, T, delt);
    cudaEventRecord(e1, T.compute);
    cuda_check(cudaEventSynchronize(e1), "sync");
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ms_max = T.allreduce((double)ms, ncclMax);
    A.synchronize();
    const double np1 = total_np();
    std::vector<double> de(A.nchunk), db(A.nchunk);
    A.check(picnix_cuda_get_diverror(A.handle(), de.data(), db.data()));
    double worst = 0;
    for (double v : de)
      worst = std::fmax(worst, std::fabs(v));
    worst = T.allreduce(worst, ncclMax);
    int64_t regrows = 0, late = 0;
    A.check(picnix_cuda_get_growth_stats(A.handle(), &regrows, &late));
    if (env.rank == 0)
      std::printf("{\"driver\": \"C++ host (host/picnix_host_nccl.hpp), NCCL send/recv overlapped with compute\", "
                  "\"n_gpus\": %d, \"cells_per_gpu\": %d, \"particles\": %.0f, \"particles_after\": %.0f, "
                  "\"steps\": %d, \"ms_per_step\": %.4f, \"particle_steps_per_s\": %.5e, "
                  "\"max_chunk_abs_sum_divE_minus_rho\": %.3e, \"late_particles\": %" PRId64 "}\n",
                  env.world, cells * cells * cells, np0, np1, steps, ms_max / steps, np0 * steps / (ms_max * 1e-3),
                  worst, late);
  } catch (const Error& e) {
    std::fprintf(stderr, "[rank %d] picnix error %d: %s\n", env.rank, e.status, e.what());
    return 1;
  }
  return 0;
}
