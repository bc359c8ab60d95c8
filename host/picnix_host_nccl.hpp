// -*- C++ -*-
// picnix_host_nccl.hpp -- the multi-rank half of the C++ host: one process per GPU, the per-peer halo
// buffers of the arena moved with NCCL point-to-point calls, one group per boundary mode, on a
// communication stream of their own so that the transfers run under the kernels of the compute stream.
//
// It plays the part of nix::Chunk::{begin,end}_bc_exchange + MPI_Isend/Irecv/Waitall
// (nix/chunk.hpp:464-543) for chunks owned by other ranks; chunks of the same rank never leave the GPU.
// The schedule of step_phases() is PicApplication::push_openmp's (pic/pic_application.cpp:219-292): the
// current and particle transfers are in flight during the second B half step, the J unpack and the E step,
// the E/B transfer during the particle unpack and the sort (the reference overlaps at the same places,
// :242-251).
//
// Plain g++: needs cuda_runtime.h and nccl.h, links -lcudart -lnccl -lpicnix_b200.  No MPI: the NCCL
// unique id travels through a file (rank 0 writes it, the others wait for it), ranks come from the
// environment a launcher such as `python -m torch.distributed.run --no-python` provides
// (RANK, WORLD_SIZE, LOCAL_RANK).
#ifndef PICNIX_HOST_NCCL_HPP
#define PICNIX_HOST_NCCL_HPP

#include <cuda_runtime.h>
#include <nccl.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <thread>

#include "picnix_host.hpp"

namespace picnix
{
namespace host
{

inline void cuda_check(cudaError_t e, const char* what)
{
  if (e != cudaSuccess)
    throw Error(PICNIX_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
inline void nccl_check(ncclResult_t r, const char* what)
{
  if (r != ncclSuccess)
    throw Error(PICNIX_ERR_CUDA, std::string(what) + ": " + ncclGetErrorString(r));
}

struct RankEnv {
  int rank = 0, world = 1, local = 0;
  RankEnv()
  {
    if (const char* e = std::getenv("RANK"))
      rank = std::atoi(e);
    if (const char* e = std::getenv("WORLD_SIZE"))
      world = std::atoi(e);
    if (const char* e = std::getenv("LOCAL_RANK"))
      local = std::atoi(e);
  }
};

class NcclTransport
{
public:
  NcclTransport(Arena& arena, const RankEnv& env, const std::string& id_file) : A(arena), env_(env)
  {
    cuda_check(cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking), "stream");
    cuda_check(cudaStreamCreateWithFlags(&comm, cudaStreamNonBlocking), "stream");
    A.check(picnix_cuda_set_stream(A.handle(), (void*)compute));
    for (int m = 0; m < 4; m++) {
      cuda_check(cudaEventCreateWithFlags(&packed[m], cudaEventDisableTiming), "event");
      cuda_check(cudaEventCreateWithFlags(&landed[m], cudaEventDisableTiming), "event");
    }
    if (env.world > 1) {
      ncclUniqueId id;
      if (env.rank == 0) {
        nccl_check(ncclGetUniqueId(&id), "ncclGetUniqueId");
        std::ofstream(id_file + ".tmp", std::ios::binary).write((const char*)&id, sizeof(id));
        std::rename((id_file + ".tmp").c_str(), id_file.c_str());
      } else {
        for (int tries = 0;; tries++) {
          std::ifstream in(id_file, std::ios::binary);
          if (in && in.read((char*)&id, sizeof(id)))
            break;
          if (tries > 6000)
            throw Error(PICNIX_ERR_INVALID, "NCCL id file never appeared: " + id_file);
          std::this_thread::sleep_for(std::chrono::milliseconds(10));
        }
      }
      nccl_check(ncclCommInitRank(&nccl, env.world, id, env.rank), "ncclCommInitRank");
      int32_t n = 0;
      A.check(picnix_cuda_get_peers(A.handle(), &n, nullptr));
      peers.resize(n);
      A.check(picnix_cuda_get_peers(A.handle(), &n, peers.data()));
      cuda_check(cudaMalloc((void**)&d_counts, sizeof(int64_t) * 2 * (n > 0 ? n : 1)), "counts");
      A.check(picnix_cuda_set_option(A.handle(), "async_migration", 1));
    }
  }
  ~NcclTransport()
  {
    cudaStreamSynchronize(compute);
    cudaStreamSynchronize(comm);
    if (nccl)
      ncclCommDestroy(nccl);
    if (d_counts)
      cudaFree(d_counts);
  }

  /// after boundary_begin(mode): the packed buffers start travelling as soon as the pack kernels are done
  void start(int mode)
  {
    if (peers.empty())
      return;
    cuda_check(cudaEventRecord(packed[mode], compute), "record");
    cuda_check(cudaStreamWaitEvent(comm, packed[mode], 0), "wait");
    if (mode == PICNIX_BOUNDARY_PARTICLE && !sizes_known())
      exchange_counts(); // first steps: exact sizes like MPI_Iprobe + MPI_Get_count (nix/chunk.cpp:329-345)
    nccl_check(ncclGroupStart(), "group");
    for (size_t i = 0; i < peers.size(); i++) {
      void *sp, *rp;
      int64_t sb, rb;
      A.check(picnix_cuda_get_comm_buffer(A.handle(), mode, (int)i, &sp, &sb, &rp, &rb));
      if (rb > 0)
        nccl_check(ncclRecv(rp, (size_t)rb, ncclChar, peers[i], nccl, comm), "recv");
      if (sb > 0)
        nccl_check(ncclSend(sp, (size_t)sb, ncclChar, peers[i], nccl, comm), "send");
    }
    nccl_check(ncclGroupEnd(), "group");
    cuda_check(cudaEventRecord(landed[mode], comm), "record");
  }

  /// before boundary_end(mode): the compute stream waits for the data of this mode
  void finish(int mode)
  {
    if (!peers.empty())
      cuda_check(cudaStreamWaitEvent(compute, landed[mode], 0), "wait");
  }

  /// max / sum over ranks of a host value (timing, particle counts)
  double allreduce(double v, ncclRedOp_t op)
  {
    if (env_.world == 1)
      return v;
    double* d = nullptr;
    cuda_check(cudaMalloc((void**)&d, sizeof(double)), "malloc");
    cuda_check(cudaMemcpyAsync(d, &v, sizeof(double), cudaMemcpyHostToDevice, comm), "copy");
    nccl_check(ncclAllReduce(d, d, 1, ncclDouble, op, nccl, comm), "allreduce");
    cuda_check(cudaMemcpyAsync(&v, d, sizeof(double), cudaMemcpyDeviceToHost, comm), "copy");
    cuda_check(cudaStreamSynchronize(comm), "sync");
    cudaFree(d);
    return v;
  }

  Arena&           A;
  cudaStream_t     compute = nullptr, comm = nullptr;
  std::vector<int> peers;

private:
  bool sizes_known()
  {
    for (size_t i = 0; i < peers.size(); i++) {
      void *sp, *rp;
      int64_t sb, rb;
      A.check(picnix_cuda_get_comm_buffer(A.handle(), PICNIX_BOUNDARY_PARTICLE, (int)i, &sp, &sb, &rp, &rb));
      if (rb == 0)
        return false;
    }
    return true;
  }
  void exchange_counts()
  {
    const size_t        n = peers.size();
    std::vector<int64_t> h(2 * n, 0);
    for (size_t i = 0; i < n; i++) {
      void *sp, *rp;
      int64_t rb;
      A.check(picnix_cuda_get_comm_buffer(A.handle(), PICNIX_BOUNDARY_PARTICLE, (int)i, &sp, &h[i], &rp, &rb));
    }
    cuda_check(cudaMemcpyAsync(d_counts, h.data(), sizeof(int64_t) * n, cudaMemcpyHostToDevice, comm), "copy");
    nccl_check(ncclGroupStart(), "group");
    for (size_t i = 0; i < n; i++) {
      nccl_check(ncclRecv(d_counts + n + i, 1, ncclInt64, peers[i], nccl, comm), "recv");
      nccl_check(ncclSend(d_counts + i, 1, ncclInt64, peers[i], nccl, comm), "send");
    }
    nccl_check(ncclGroupEnd(), "group");
    cuda_check(cudaMemcpyAsync(h.data() + n, d_counts + n, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, comm), "copy");
    cuda_check(cudaStreamSynchronize(comm), "sync");
    for (size_t i = 0; i < n; i++)
      A.check(picnix_cuda_set_recv_bytes(A.handle(), PICNIX_BOUNDARY_PARTICLE, (int)i, h[n + i]));
  }

  RankEnv     env_;
  ncclComm_t  nccl = nullptr;
  cudaEvent_t packed[4], landed[4];
  int64_t*    d_counts = nullptr;
};

/// one full boundary exchange (set-up, diagnostics)
inline void exchange(Arena& A, NcclTransport& T, int mode)
{
  A.check(picnix_cuda_boundary_begin(A.handle(), mode));
  T.start(mode);
  T.finish(mode);
  A.check(picnix_cuda_boundary_end(A.handle(), mode));
}

/// One time step in the order of PicApplication::push_openmp (pic/pic_application.cpp:219-292) for all
/// chunks of this rank, transfers overlapped with the kernels between start() and finish().
inline void step_phases(Arena& A, NcclTransport& T, double delt)
{
  picnix_arena_t* h = A.handle();
  A.check(picnix_cuda_push_bfd(h, 0, -1, 0.5 * delt));
  A.check(picnix_cuda_push_deposit_fused(h, 0, -1, delt));
  A.check(picnix_cuda_boundary_begin(h, PICNIX_BOUNDARY_CUR));
  T.start(PICNIX_BOUNDARY_CUR);
  A.check(picnix_cuda_boundary_begin(h, PICNIX_BOUNDARY_PARTICLE));
  T.start(PICNIX_BOUNDARY_PARTICLE);
  A.check(picnix_cuda_push_bfd(h, 0, -1, 0.5 * delt));
  T.finish(PICNIX_BOUNDARY_CUR);
  A.check(picnix_cuda_boundary_end(h, PICNIX_BOUNDARY_CUR));
  A.check(picnix_cuda_push_efd(h, 0, -1, delt));
  A.check(picnix_cuda_boundary_begin(h, PICNIX_BOUNDARY_EMF));
  T.start(PICNIX_BOUNDARY_EMF);
  T.finish(PICNIX_BOUNDARY_PARTICLE);
  A.check(picnix_cuda_boundary_end(h, PICNIX_BOUNDARY_PARTICLE));
  T.finish(PICNIX_BOUNDARY_EMF);
  A.check(picnix_cuda_boundary_end(h, PICNIX_BOUNDARY_EMF));
}

} // namespace host
} // namespace picnix

#endif
