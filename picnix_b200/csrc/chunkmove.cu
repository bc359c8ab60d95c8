// -*- C++ -*-
// Chunk moves between ranks (load rebalancing): the complete state of one local chunk as ONE
// contiguous DEVICE buffer.
//
// The reference moves chunks between neighbouring ranks after Balancer::assign changed the rank
// boundaries (nix/application.hpp:332 rebalance, nix/balancer.hpp:122-332) by PicChunk::pack ->
// MPI -> PicChunk::unpack on host buffers (pic/pic_chunk.cpp:59-95).  Here the packed chunk stays
// on the device, so the caller hands it to NCCL / a peer copy directly:
//
//   int64  header[4 + Ns]   magic, Ng, Ns, reserved, np[Ns]
//   f64    uf[Ng][6], uj[Ng][4], ff[Ng][3][3]            (device layouts)
//   f64    particles: per species [7][np] (structure of arrays, component major)
//
// The receiving arena must have been given segment capacities >= np (the caller learns np from
// picnix_cuda_get_np on the sending side).  Unpacked particles keep their order; the caller runs
// picnix_cuda_sort_particle once all chunks have arrived (keys and pindex are rebuilt there).
#include "arena.hpp"

namespace picnix
{
namespace
{
constexpr int64_t CHUNK_MAGIC = 0x50494e43484b3031ll; // "PINCHK01"

int64_t packed_bytes(const picnix_arena* a, const int* np)
{
  const Geom& g = a->g;
  int64_t     n = (4 + g.Ns) * (int64_t)sizeof(int64_t) + (int64_t)g.Ng * (6 + 4 + 9) * sizeof(double);
  for (int is = 0; is < g.Ns; is++)
    n += (int64_t)np[is] * NC * sizeof(double);
  return n;
}
} // namespace
} // namespace picnix

using namespace picnix;

extern "C" {

int picnix_cuda_chunk_pack_size(picnix_arena_t* a, int32_t ichunk, int64_t* bytes)
{
  if (a == nullptr || bytes == nullptr || ichunk < 0 || ichunk >= a->g.nchunk)
    return PICNIX_ERR_INVALID;
  if (!a->particles_allocated)
    return fail(a, PICNIX_ERR_INVALID, "no particles allocated");
  std::vector<int> np(a->g.Ns);
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  PICNIX_CUDA(a, cudaMemcpy(np.data(), a->d.np + (size_t)ichunk * a->g.Ns, a->g.Ns * sizeof(int),
                            cudaMemcpyDeviceToHost));
  *bytes = packed_bytes(a, np.data());
  return PICNIX_OK;
}

int picnix_cuda_chunk_pack(picnix_arena_t* a, int32_t ichunk, void* dev_buf, int64_t bytes)
{
  if (a == nullptr || dev_buf == nullptr || ichunk < 0 || ichunk >= a->g.nchunk)
    return PICNIX_ERR_INVALID;
  if (!a->particles_allocated)
    return fail(a, PICNIX_ERR_INVALID, "no particles allocated");
  const Geom&      g = a->g;
  std::vector<int> np(g.Ns);
  {
    int mstatus = materialize_sort(a);
    if (mstatus != PICNIX_OK)
      return mstatus;
  }
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  PICNIX_CUDA(a, cudaMemcpy(np.data(), a->d.np + (size_t)ichunk * g.Ns, g.Ns * sizeof(int),
                            cudaMemcpyDeviceToHost));
  if (bytes < packed_bytes(a, np.data()))
    return fail(a, PICNIX_ERR_OVERFLOW, "chunk_pack: buffer too small");

  std::vector<int64_t> header(4 + g.Ns, 0);
  header[0] = CHUNK_MAGIC;
  header[1] = g.Ng;
  header[2] = g.Ns;
  for (int is = 0; is < g.Ns; is++)
    header[4 + is] = np[is];
  char* out = static_cast<char*>(dev_buf);
  PICNIX_CUDA(a, cudaMemcpyAsync(out, header.data(), header.size() * sizeof(int64_t),
                                 cudaMemcpyHostToDevice, a->stream));
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream)); // header is a stack-lifetime host vector
  out += header.size() * sizeof(int64_t);
  const size_t nf[3]  = {(size_t)g.Ng * 6, (size_t)g.Ng * 4, (size_t)g.Ng * 9};
  double*      src[3] = {a->d.uf + (size_t)ichunk * nf[0], a->d.uj + (size_t)ichunk * nf[1],
                         a->d.ff + (size_t)ichunk * nf[2]};
  for (int f = 0; f < 3; f++) {
    PICNIX_CUDA(a, cudaMemcpyAsync(out, src[f], nf[f] * sizeof(double), cudaMemcpyDeviceToDevice, a->stream));
    out += nf[f] * sizeof(double);
  }
  for (int is = 0; is < g.Ns; is++) {
    const int64_t off = a->seg_off[(size_t)ichunk * g.Ns + is];
    for (int k = 0; k < NC; k++) {
      if (np[is] > 0)
        PICNIX_CUDA(a, cudaMemcpyAsync(out, a->d.xu + k * a->d.pcap + off, (size_t)np[is] * sizeof(double),
                                       cudaMemcpyDeviceToDevice, a->stream));
      out += (size_t)np[is] * sizeof(double);
    }
  }
  return check_cuda(a, cudaStreamSynchronize(a->stream), "chunk_pack");
}

int picnix_cuda_chunk_unpack(picnix_arena_t* a, int32_t ichunk, const void* dev_buf, int64_t bytes)
{
  if (a == nullptr || dev_buf == nullptr || ichunk < 0 || ichunk >= a->g.nchunk)
    return PICNIX_ERR_INVALID;
  if (!a->particles_allocated)
    return fail(a, PICNIX_ERR_INVALID, "chunk_unpack: set_particle_capacity first");
  {
    int mstatus = materialize_sort(a);
    if (mstatus != PICNIX_OK)
      return mstatus;
  }
  const Geom&          g = a->g;
  std::vector<int64_t> header(4 + g.Ns, 0);
  if (bytes < (int64_t)(header.size() * sizeof(int64_t)))
    return fail(a, PICNIX_ERR_INVALID, "chunk_unpack: truncated buffer");
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  PICNIX_CUDA(a, cudaMemcpy(header.data(), dev_buf, header.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
  if (header[0] != CHUNK_MAGIC || header[1] != g.Ng || header[2] != g.Ns)
    return fail(a, PICNIX_ERR_INVALID, "chunk_unpack: buffer does not hold a chunk of this geometry");
  std::vector<int> np(g.Ns);
  for (int is = 0; is < g.Ns; is++) {
    np[is] = (int)header[4 + is];
    if (np[is] < 0 || np[is] > a->seg_cap[(size_t)ichunk * g.Ns + is])
      return fail(a, PICNIX_ERR_OVERFLOW, "chunk_unpack: np exceeds segment capacity");
  }
  if (bytes < packed_bytes(a, np.data()))
    return fail(a, PICNIX_ERR_INVALID, "chunk_unpack: truncated buffer");

  const char*  in     = static_cast<const char*>(dev_buf) + header.size() * sizeof(int64_t);
  const size_t nf[3]  = {(size_t)g.Ng * 6, (size_t)g.Ng * 4, (size_t)g.Ng * 9};
  double*      dst[3] = {a->d.uf + (size_t)ichunk * nf[0], a->d.uj + (size_t)ichunk * nf[1],
                         a->d.ff + (size_t)ichunk * nf[2]};
  for (int f = 0; f < 3; f++) {
    PICNIX_CUDA(a, cudaMemcpyAsync(dst[f], in, nf[f] * sizeof(double), cudaMemcpyDeviceToDevice, a->stream));
    in += nf[f] * sizeof(double);
  }
  for (int is = 0; is < g.Ns; is++) {
    const int64_t off = a->seg_off[(size_t)ichunk * g.Ns + is];
    for (int k = 0; k < NC; k++) {
      if (np[is] > 0)
        PICNIX_CUDA(a, cudaMemcpyAsync(a->d.xu + k * a->d.pcap + off, in, (size_t)np[is] * sizeof(double),
                                       cudaMemcpyDeviceToDevice, a->stream));
      in += (size_t)np[is] * sizeof(double);
    }
  }
  PICNIX_CUDA(a, cudaMemcpyAsync(a->d.np + (size_t)ichunk * g.Ns, np.data(), g.Ns * sizeof(int),
                                 cudaMemcpyHostToDevice, a->stream));
  PICNIX_CUDA(a, cudaMemsetAsync(a->d.ntail + (size_t)ichunk * g.Ns, 0, g.Ns * sizeof(int), a->stream));
  a->pindex_valid     = false;
  a->leave_list_valid = false;
  return check_cuda(a, cudaStreamSynchronize(a->stream), "chunk_unpack");
}

} // extern "C"
