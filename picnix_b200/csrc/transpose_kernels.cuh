// -*- C++ -*-
// Layout-conversion kernels at the host boundary, shared by arena.cu (per-chunk transfers) and
// hostio.cu (pipelined whole-arena transfers).  `static`: every translation unit that launches
// them gets its own copy (the library is built without relocatable device code).
#ifndef PICNIX_B200_TRANSPOSE_KERNELS_CUH
#define PICNIX_B200_TRANSPOSE_KERNELS_CUH

#include "arena.hpp"

namespace picnix
{

//
// AoS <-> SoA transposes at the host boundary (the reference's particle array is [Np][7])
//
static __global__ void aos_to_soa_kernel(const double* __restrict__ aos, double* __restrict__ soa,
                                  int64_t off, int64_t pcap, int n)
{
  // one thread per (particle, component) of the staged AoS block: coalesced reads, strided writes
  // that still fall in 7 contiguous runs per warp
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * NC)
    return;
  int ip            = (int)(i / NC);
  int ic            = (int)(i - (int64_t)ip * NC);
  soa[ic * pcap + off + ip] = aos[i];
}

static __global__ void soa_to_aos_kernel(const double* __restrict__ soa, double* __restrict__ aos,
                                  int64_t off, int64_t pcap, int n)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * NC)
    return;
  int ip = (int)(i / NC);
  int ic = (int)(i - (int64_t)ip * NC);
  aos[i] = soa[ic * pcap + off + ip];
}

// Many small segments (1-D runs: tens of thousands of 512-particle segments): one launch for a GROUP of
// consecutive segments whose host span was copied in one piece.  capoff = prefix sums of the host
// capacities (the host array keeps its capacity gaps, so does the staged copy); blockIdx.y = segment.
static __global__ void aos_to_soa_group_kernel(const double* __restrict__ span, double* __restrict__ soa,
                                               const int64_t* __restrict__ capoff,
                                               const int64_t* __restrict__ seg_off,
                                               const int* __restrict__ np, int s0, int64_t pcap)
{
  const int     s = s0 + blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)np[s] * NC)
    return;
  const int ip = (int)(i / NC);
  const int ic = (int)(i - (int64_t)ip * NC);
  soa[ic * pcap + seg_off[s] + ip] = span[(capoff[s] - capoff[s0]) * NC + i];
}

static __global__ void soa_to_aos_group_kernel(const double* __restrict__ soa, double* __restrict__ span,
                                               const int64_t* __restrict__ capoff,
                                               const int64_t* __restrict__ seg_off,
                                               const int* __restrict__ np, int s0, int64_t pcap)
{
  const int     s = s0 + blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)np[s] * NC)
    return;
  const int ip = (int)(i / NC);
  const int ic = (int)(i - (int64_t)ip * NC);
  span[(capoff[s] - capoff[s0]) * NC + i] = soa[ic * pcap + seg_off[s] + ip];
}

//
// ff: host layout [cell][3][6] (reference) <-> device layout [cell][3][3]
//
static __global__ void ff_expand_kernel(const double* __restrict__ dev, double* __restrict__ host_layout,
                                 int64_t ncell)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncell * 18)
    return;
  int64_t cell = i / 18;
  int     r    = (int)(i - cell * 18);
  int     t    = r / 6;
  int     k    = r - t * 6;
  host_layout[i] = k < 3 ? dev[cell * 9 + t * 3 + k] : 0.0;
}

static __global__ void ff_compact_kernel(const double* __restrict__ host_layout, double* __restrict__ dev,
                                  int64_t ncell)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncell * 9)
    return;
  int64_t cell = i / 9;
  int     r    = (int)(i - cell * 9);
  int     t    = r / 3;
  int     k    = r - t * 3;
  dev[i]       = host_layout[cell * 18 + t * 6 + k];
}


} // namespace picnix

#endif
