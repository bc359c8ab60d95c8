// -*- C++ -*-
// Device routines shared by the particle exchange (halo.cu) and the segment growth (grow.cu).
#ifndef PICNIX_B200_MIGRATE_CUH
#define PICNIX_B200_MIGRATE_CUH

#include "particle_common.cuh"

namespace picnix
{

// A migrant whose destination is full -- a segment, or the bounded message to a peer -- is kept in the
// spill list instead of being dropped: tag.x >= 0 is the destination segment (re-appended by grow.cu once
// the segment has grown), tag.x < 0 is ~(message slot) of a remote neighbour (re-sent with the next
// exchange).  Only a full spill list loses particles, and that is an error.
__device__ __forceinline__ void spill_particle(const DevPtrs& d, const double* p, int dest, int tagy)
{
  const int k = atomicAdd(d.spill_count, 1);
  if (k >= d.spill_cap) {
    atomicExch(d.errflag + 0, 1);
    return;
  }
  double* out = d.spill_rec + (int64_t)k * 8;
#pragma unroll
  for (int c = 0; c < NC; c++)
    out[c] = p[c];
  int2 tag = make_int2(dest, tagy);
  out[7]   = *reinterpret_cast<double*>(&tag);
}

// append one particle behind the active particles of (chunk, species); p = 7 components
__device__ __forceinline__ void append_particle(const Geom& g, const DevPtrs& d, int chunk, int is,
                                                double* p)
{
  const int seg  = chunk * g.Ns + is;
  const int slot = atomicAdd(d.ntail + seg, 1);
  const int ip   = d.np[seg] + slot;
  if (ip >= d.seg_cap[seg]) {
    // XtensorHaloParticle3D::pre_unpack would have resized first (nix/xtensor_halo3d.hpp:406-418): wait
    // in the spill list for resolve_growth()
    atomicSub(d.ntail + seg, 1);
    spill_particle(d, p, seg, is);
    return;
  }
  // post_unpack: periodic wrap, then count in the receiving chunk's geometry
  wrap_periodic(g, p[0], p[1], p[2]);
  const int64_t i = d.seg_off[seg] + ip;
#pragma unroll
  for (int k = 0; k < NC; k++)
    d.xu[k * d.pcap + i] = p[k];
  const int key = cell_key(g, d.clim + chunk * 6, p[0], p[1], p[2]);
  d.gindex[i]   = key;
  atomicAdd(d.pcount + (int64_t)seg * (g.Ng + 1) + key, 1);
}

} // namespace picnix

#endif
