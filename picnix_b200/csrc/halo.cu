// -*- C++ -*-
// K5: boundary exchange between chunks -- fields (copy), currents (add) and particle migration.
//
// Reference: nix::Chunk::{pack,begin,probe,end,unpack}_bc_exchange (nix/chunk.hpp:392-543,
// nix/chunk.cpp:310-395) with the policies XtensorHaloField3D / XtensorHaloCurrent3D /
// XtensorHaloParticle3D (nix/xtensor_halo3d.hpp:18-129, 192-499).  There every one of the 26
// neighbour relations is an MPI message, even between chunks of the same rank.  Here:
//   * neighbours inside the arena need NO message: one gather kernel reads the neighbour's cells
//     directly (field: interior margin -> ghost, nix/chunk.cpp:171-207; current: ghost -> += interior
//     margin, nix/xtensor_halo3d.hpp:93-99,119-125), visiting the directions in the reference's
//     unpack order (dirz, diry, dirx ascending) so the floating-point sum order is the same;
//   * chunks owned by another rank are served through ONE contiguous send and ONE receive buffer
//     per peer rank and mode; the caller moves send -> recv (NCCL / P2P) between begin and end.
//     Messages inside a buffer are ordered by (sender chunk id, sender direction) on both sides.
//   * particles leaving a chunk are appended straight behind the destination chunk's active
//     particles (periodic wrap + cell key + histogram update included, i.e. post_unpack's work,
//     nix/xtensor_halo3d.hpp:477-491), or into the peer's staging buffer as 64-byte records.
#include "migrate.cuh"

#include <algorithm>
#include <tuple>

namespace picnix
{

namespace
{

constexpr int HALO_THREADS = 128;

// extent of the halo region of direction code dcode (0,1,2 per axis) along one axis
__host__ __device__ inline int region_len(const Geom& g, int axis, int dcode)
{
  return dcode == 1 ? (g.Ub[axis] - g.Lb[axis] + 1) : g.nb;
}

// first index of the interior-margin region ("send_bound" of the field halo)
__host__ __device__ inline int margin_lo(const Geom& g, int axis, int dcode)
{
  return dcode == 2 ? g.Ub[axis] - g.nb + 1 : g.Lb[axis];
}

// first index of the ghost region ("recv_bound")
__host__ __device__ inline int ghost_lo(const Geom& g, int axis, int dcode)
{
  return dcode == 0 ? g.Lb[axis] - g.nb : (dcode == 1 ? g.Lb[axis] : g.Ub[axis] + 1);
}

__host__ __device__ inline bool dir_active(const Geom& g, int dz, int dy, int dx)
{
  // ignorable dimensions only take part with direction 0 (nix/chunk.cpp:141-169)
  if (dz == 1 && dy == 1 && dx == 1)
    return false;
  if (!g.has_dim[0] && dz != 1)
    return false;
  if (!g.has_dim[1] && dy != 1)
    return false;
  if (!g.has_dim[2] && dx != 1)
    return false;
  return true;
}

// ---------------------------------------------------------------------------------------------
// field halo, local neighbours: every ghost cell pulls from the owning neighbour's interior
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HALO_THREADS) field_halo_local_kernel(Geom g, DevPtrs d)
{
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)g.nchunk * g.Ng)
    return;
  int chunk = (int)(t / g.Ng);
  int r     = (int)(t - (int64_t)chunk * g.Ng);
  int idx[3];
  idx[0] = r / (g.M[1] * g.M[2]);
  r -= idx[0] * g.M[1] * g.M[2];
  idx[1] = r / g.M[2];
  idx[2] = r - idx[1] * g.M[2];

  int dcode[3], src[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    if (!g.has_dim[a]) {
      if (idx[a] != g.Lb[a])
        return; // ghost planes of ignorable dimensions are never exchanged
      dcode[a] = 1;
      src[a]   = idx[a];
    } else if (idx[a] < g.Lb[a]) {
      dcode[a] = 0;
      src[a]   = idx[a] + g.dims[a];
    } else if (idx[a] > g.Ub[a]) {
      dcode[a] = 2;
      src[a]   = idx[a] - g.dims[a];
    } else {
      dcode[a] = 1;
      src[a]   = idx[a];
    }
  }
  if (dcode[0] == 1 && dcode[1] == 1 && dcode[2] == 1)
    return;

  int nb = d.nbr[chunk * NBSIZE + 9 * dcode[0] + 3 * dcode[1] + dcode[2]];
  if (nb < 0)
    return; // none, or remote (served from the receive buffer in halo_end)

  int64_t s = (((int64_t)nb * g.M[0] + src[0]) * g.M[1] + src[1]) * g.M[2] + src[2];
#pragma unroll
  for (int k = 0; k < 6; k++)
    d.uf[t * 6 + k] = d.uf[s * 6 + k];
}

// ---------------------------------------------------------------------------------------------
// current halo, local neighbours: every interior-margin cell adds the neighbours' ghost cells
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HALO_THREADS) current_halo_local_kernel(Geom g, DevPtrs d)
{
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)g.nchunk * g.Ng)
    return;
  int chunk = (int)(t / g.Ng);
  int r     = (int)(t - (int64_t)chunk * g.Ng);
  int idx[3];
  idx[0] = r / (g.M[1] * g.M[2]);
  r -= idx[0] * g.M[1] * g.M[2];
  idx[1] = r / g.M[2];
  idx[2] = r - idx[1] * g.M[2];

  // only interior cells receive; for each axis, which direction codes reach this cell
  bool reach[3][3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    if (idx[a] < g.Lb[a] || idx[a] > g.Ub[a])
      return;
    reach[a][0] = g.has_dim[a] && idx[a] < g.Lb[a] + g.nb;
    reach[a][1] = true;
    reach[a][2] = g.has_dim[a] && idx[a] > g.Ub[a] - g.nb;
  }

  double acc[4];
#pragma unroll
  for (int k = 0; k < 4; k++)
    acc[k] = d.uj[t * 4 + k];
  bool touched = false;

  for (int dz = 0; dz < 3; dz++) {
    if (!reach[0][dz])
      continue;
    for (int dy = 0; dy < 3; dy++) {
      if (!reach[1][dy])
        continue;
      for (int dx = 0; dx < 3; dx++) {
        if (!reach[2][dx] || (dz == 1 && dy == 1 && dx == 1))
          continue;
        int nb = d.nbr[chunk * NBSIZE + 9 * dz + 3 * dy + dx];
        if (nb < 0)
          continue;
        // the neighbour's ghost cell that overlaps this interior cell
        int sz = idx[0] + (dz == 0 ? g.dims[0] : (dz == 2 ? -g.dims[0] : 0));
        int sy = idx[1] + (dy == 0 ? g.dims[1] : (dy == 2 ? -g.dims[1] : 0));
        int sx = idx[2] + (dx == 0 ? g.dims[2] : (dx == 2 ? -g.dims[2] : 0));
        int64_t s = (((int64_t)nb * g.M[0] + sz) * g.M[1] + sy) * g.M[2] + sx;
#pragma unroll
        for (int k = 0; k < 4; k++)
          acc[k] += d.uj[s * 4 + k];
        touched = true;
      }
    }
  }
  if (touched) {
#pragma unroll
    for (int k = 0; k < 4; k++)
      d.uj[t * 4 + k] = acc[k];
  }
}

// ---------------------------------------------------------------------------------------------
// remote messages for the fixed-size modes.  desc[m] = (local chunk, direction code 0..26)
//   pack  : EMF sends the interior margin, CUR sends the ghost region
//   unpack: EMF stores into the ghost region, CUR adds into the interior margin
// One block per message; the element order inside a message is (z, y, x, component) like the
// reference's strided_view copy.
// ---------------------------------------------------------------------------------------------
template <int NCOMP, bool IS_CURRENT, bool IS_PACK>
__global__ void __launch_bounds__(HALO_THREADS)
remote_halo_kernel(Geom g, double* __restrict__ field, const int* __restrict__ desc,
                   const int64_t* __restrict__ msg_off, double* __restrict__ buf, int ncomp_rt)
{
  const int ncomp = NCOMP > 0 ? NCOMP : ncomp_rt; // NCOMP == 0: moments, Ns * 14 components
  const int m     = blockIdx.x;
  const int chunk = desc[2 * m + 0];
  const int dir   = desc[2 * m + 1];
  const int dc[3] = {dir / 9, (dir / 3) % 3, dir % 3};

  int lo[3], len[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    len[a] = region_len(g, a, dc[a]);
    // pack/EMF and unpack/CUR touch the interior margin; pack/CUR and unpack/EMF the ghosts
    bool interior = (IS_PACK != IS_CURRENT);
    lo[a]         = interior ? margin_lo(g, a, dc[a]) : ghost_lo(g, a, dc[a]);
  }
  const int n   = len[0] * len[1] * len[2] * ncomp;
  double*   msg = buf + msg_off[m];

  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    int k    = e % ncomp;
    int cell = e / ncomp;
    int jx   = cell % len[2];
    int jy   = (cell / len[2]) % len[1];
    int jz   = cell / (len[2] * len[1]);
    int64_t c =
        (((int64_t)chunk * g.M[0] + lo[0] + jz) * g.M[1] + lo[1] + jy) * g.M[2] + lo[2] + jx;
    if (IS_PACK) {
      msg[e] = field[c * ncomp + k];
    } else if (IS_CURRENT) {
      // the margins of a face, its edges and its corners overlap and every message is its own
      // thread block: the add must be atomic (a plain += loses updates between messages)
      atomicAdd(field + c * ncomp + k, msg[e]);
    } else {
      field[c * ncomp + k] = msg[e];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// particle migration
// ---------------------------------------------------------------------------------------------
struct MigrateTables {
  const int* slot_peer;  // remote slot -> peer index
  double**   psend;      // [npeer] staging buffers (records of 8 doubles)
  int**      psend_cnt;  // [npeer] record counters
  const int64_t* psend_cap; // [npeer] capacity in records
  const int* nbid;       // [nchunk][27] global neighbour ids (for remote records)
};

// send one particle that left (chunk, species) to the neighbour its position points at
__device__ __forceinline__ void migrate_particle(const Geom& g, const DevPtrs& d,
                                                 const MigrateTables& tab, int seg, int64_t i)
{
  const int chunk = seg / g.Ns;
  const int is    = seg - chunk * g.Ns;
  double    p[NC];
#pragma unroll
  for (int k = 0; k < NC; k++)
    p[k] = d.xu[k * d.pcap + i];

  const int dir = direction_code(g, d.clim + chunk * 6, p[0], p[1], p[2]);
  if (dir == 13)
    return;
  const int nb = d.nbr[chunk * NBSIZE + dir];
  if (nb >= 0) {
    append_particle(g, d, nb, is, p);
  } else if (nb <= NB_REMOTE_BASE) {
    const int     slot = NB_REMOTE_BASE - nb;
    const int     peer = tab.slot_peer[slot];
    const int     rec  = atomicAdd(tab.psend_cnt[peer], 1);
    if (rec >= tab.psend_cap[peer]) {
      // the message to this peer is full (lagged-count bound): send it with the next exchange
      atomicSub(tab.psend_cnt[peer], 1);
      spill_particle(d, p, ~slot, tab.nbid[chunk * NBSIZE + dir] | (is << 24));
      atomicExch(d.errflag + 1, 1); // reported as late delivery, not as an error
      return;
    }
    double* out = tab.psend[peer] + (int64_t)rec * 8;
#pragma unroll
    for (int k = 0; k < NC; k++)
      out[k] = p[k];
    // destination: global chunk id and species, packed into the eighth slot
    int2 tag = make_int2(tab.nbid[chunk * NBSIZE + dir], is);
    out[7]   = *reinterpret_cast<double*>(&tag);
  }
  // NB_NONE: open boundary, the particle is simply dropped by the sort (MPI_PROC_NULL send)
}

// Leavers listed by the fused push kernel: one thread per list entry (about 1 % of the particles).
// An overflowed list is left to the scan kernel below.
__global__ void __launch_bounds__(HALO_THREADS)
migrate_list_kernel(Geom g, DevPtrs d, MigrateTables tab)
{
  const int n = *d.leave_count;
  if (n > d.leave_cap)
    return;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int64_t e = d.leave_idx[k];
    migrate_particle(g, d, tab, (int)(e >> 40), e & (((int64_t)1 << 40) - 1));
  }
}

// Scan of all keys (XtensorHaloParticle3D::pre_pack looks at every particle too,
// nix/xtensor_halo3d.hpp:192-260).  Used when no leaver list describes the current keys
// (use_list == 0), and as the fallback of an overflowed list (use_list == 1).
__global__ void __launch_bounds__(HALO_THREADS)
migrate_kernel(Geom g, DevPtrs d, MigrateTables tab, int stride, int use_list)
{
  if (use_list && *d.leave_count <= d.leave_cap)
    return;
  const int64_t total = (int64_t)g.nchunk * g.Ns * stride;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int seg = (int)(idx / stride);
    const int ip  = (int)(idx - (int64_t)seg * stride);
    if (ip >= d.np[seg])
      continue;
    const int64_t i = d.seg_off[seg] + ip;
    if (d.gindex[i] != g.Ng)
      continue; // still inside its chunk
    migrate_particle(g, d, tab, seg, i);
  }
}

__global__ void __launch_bounds__(HALO_THREADS)
unpack_particle_kernel(Geom g, DevPtrs d, const double* __restrict__ recv, int nrec,
                       int chunk_begin)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrec)
    return;
  const double* in = recv + (int64_t)r * 8;
  double        p[NC];
#pragma unroll
  for (int k = 0; k < NC; k++)
    p[k] = in[k];
  double tagbits = in[7];
  int2   tag     = *reinterpret_cast<int2*>(&tagbits);
  int    chunk   = tag.x - chunk_begin;
  if (chunk < 0 || chunk >= g.nchunk || tag.y < 0 || tag.y >= g.Ns) {
    atomicExch(d.errflag + 2, 1);
    return;
  }
  append_particle(g, d, chunk, tag.y, p);
}

int64_t region_elems(const Geom& g, int dir, int ncomp)
{
  int dc[3] = {dir / 9, (dir / 3) % 3, dir % 3};
  return (int64_t)region_len(g, 0, dc[0]) * region_len(g, 1, dc[1]) * region_len(g, 2, dc[2]) *
         ncomp;
}

template <typename T>
int upload_vector(picnix_arena* a, T** dptr, const std::vector<T>& host)
{
  PICNIX_CUDA(a, cudaMalloc((void**)dptr, std::max<size_t>(host.size(), 1) * sizeof(T)));
  if (!host.empty())
    PICNIX_CUDA(a, cudaMemcpy(*dptr, host.data(), host.size() * sizeof(T),
                              cudaMemcpyHostToDevice));
  return PICNIX_OK;
}

} // namespace

// Neighbour codes + per-peer message plan.  Called once from arena_create.
int build_comm_plan(picnix_arena* a)
{
  const Geom& g      = a->g;
  const int   myrank = a->cfg.rank;

  struct Msg {
    int sender_id, sender_dir, chunk, dir;
  };
  std::vector<std::vector<Msg>> send_by_rank(a->cfg.nrank), recv_by_rank(a->cfg.nrank);

  for (int ic = 0; ic < g.nchunk; ic++) {
    for (int dz = 0; dz < 3; dz++) {
      for (int dy = 0; dy < 3; dy++) {
        for (int dx = 0; dx < 3; dx++) {
          int k    = 9 * dz + 3 * dy + dx;
          int nbid = a->nbid[(size_t)ic * NBSIZE + k];
          int nbrk = a->nbrank[(size_t)ic * NBSIZE + k];
          int code = NB_NONE;
          if (nbid >= 0 && dir_active(g, dz, dy, dx)) {
            if (nbrk == myrank) {
              code = nbid - a->chunk_begin;
            } else {
              // provisional: remote, slot assigned below
              code = NB_REMOTE_BASE;
              int opp = 9 * (2 - dz) + 3 * (2 - dy) + (2 - dx);
              send_by_rank[nbrk].push_back({a->chunk_begin + ic, k, ic, k});
              recv_by_rank[nbrk].push_back({nbid, opp, ic, k});
            }
          }
          a->nbr_code[(size_t)ic * NBSIZE + k] = code;
        }
      }
    }
  }

  auto by_sender = [](const Msg& x, const Msg& y) {
    return std::tie(x.sender_id, x.sender_dir) < std::tie(y.sender_id, y.sender_dir);
  };

  a->peers.clear();
  a->slot_peer.clear();
  for (int r = 0; r < a->cfg.nrank; r++) {
    if (send_by_rank[r].empty() && recv_by_rank[r].empty())
      continue;
    std::sort(send_by_rank[r].begin(), send_by_rank[r].end(), by_sender);
    std::sort(recv_by_rank[r].begin(), recv_by_rank[r].end(), by_sender);

    PeerPlan p;
    p.rank = r;
    for (int mode = 0; mode < 3; mode++) {
      int     ncomp = mode == 0 ? 6 : (mode == 1 ? 4 : g.Ns * 14);
      int64_t off   = 0;
      for (auto& m : send_by_rank[r]) {
        p.send_msg_off[mode].push_back(off);
        off += region_elems(g, m.dir, ncomp);
      }
      p.send_elems[mode] = off;
      off                = 0;
      for (auto& m : recv_by_rank[r]) {
        p.recv_msg_off[mode].push_back(off);
        off += region_elems(g, m.dir, ncomp);
      }
      p.recv_elems[mode] = off;
    }
    int peer_index = (int)a->peers.size();
    for (auto& m : send_by_rank[r]) {
      p.send_chunk.push_back(m.chunk);
      p.send_dir.push_back(m.dir);
      // one remote slot per (chunk, dir); the slot only needs to identify the peer
      int slot = (int)a->slot_peer.size();
      a->slot_peer.push_back(peer_index);
      a->nbr_code[(size_t)m.chunk * NBSIZE + m.dir] = NB_REMOTE_BASE - slot;
    }
    for (auto& m : recv_by_rank[r]) {
      p.recv_chunk.push_back(m.chunk);
      p.recv_dir.push_back(m.dir);
    }
    a->peers.push_back(std::move(p));
  }

  PICNIX_CUDA(a, cudaMemcpy(a->d.nbr, a->nbr_code.data(), a->nbr_code.size() * sizeof(int),
                            cudaMemcpyHostToDevice));

  // device side of the plan
  int status;
  // one allocation per fixed-size mode and direction; a peer's buffer is a 16-byte aligned slice of it
  std::vector<int64_t> sbase[2], rbase[2];
  for (int mode = 0; mode < 2; mode++) {
    int64_t stot = 0, rtot = 0;
    for (auto& p : a->peers) {
      sbase[mode].push_back(stot);
      rbase[mode].push_back(rtot);
      stot += (std::max<int64_t>(p.send_elems[mode], 1) + 1) & ~(int64_t)1;
      rtot += (std::max<int64_t>(p.recv_elems[mode], 1) + 1) & ~(int64_t)1;
    }
    if (!a->peers.empty()) {
      PICNIX_CUDA(a, cudaMalloc((void**)&a->d_send_all[mode], stot * sizeof(double)));
      PICNIX_CUDA(a, cudaMalloc((void**)&a->d_recv_all[mode], rtot * sizeof(double)));
    }
  }
  std::vector<int>     sdesc_all, rdesc_all;
  std::vector<int64_t> soff_all[2], roff_all[2];
  for (size_t ip = 0; ip < a->peers.size(); ip++) {
    PeerPlan&        p = a->peers[ip];
    std::vector<int> sdesc, rdesc;
    for (size_t m = 0; m < p.send_chunk.size(); m++) {
      sdesc.push_back(p.send_chunk[m]);
      sdesc.push_back(p.send_dir[m]);
    }
    for (size_t m = 0; m < p.recv_chunk.size(); m++) {
      rdesc.push_back(p.recv_chunk[m]);
      rdesc.push_back(p.recv_dir[m]);
    }
    sdesc_all.insert(sdesc_all.end(), sdesc.begin(), sdesc.end());
    rdesc_all.insert(rdesc_all.end(), rdesc.begin(), rdesc.end());
    if ((status = upload_vector(a, &p.d_send_desc, sdesc)) != PICNIX_OK)
      return status;
    if ((status = upload_vector(a, &p.d_recv_desc, rdesc)) != PICNIX_OK)
      return status;
    for (int mode = 0; mode < 3; mode++) {
      if ((status = upload_vector(a, &p.d_send_off[mode], p.send_msg_off[mode])) != PICNIX_OK)
        return status;
      if ((status = upload_vector(a, &p.d_recv_off[mode], p.recv_msg_off[mode])) != PICNIX_OK)
        return status;
      if (mode == PICNIX_BOUNDARY_MOM)
        continue; // diagnostics cadence: buffers are allocated by the first moment exchange
      p.d_send[mode] = a->d_send_all[mode] + sbase[mode][ip];
      p.d_recv[mode] = a->d_recv_all[mode] + rbase[mode][ip];
      for (int64_t o : p.send_msg_off[mode])
        soff_all[mode].push_back(o + sbase[mode][ip]);
      for (int64_t o : p.recv_msg_off[mode])
        roff_all[mode].push_back(o + rbase[mode][ip]);
    }
  }
  // migration counters of all peers in one array: one clear and one copy to the host per step
  if (!a->peers.empty()) {
    PICNIX_CUDA(a, cudaMalloc((void**)&a->d_mig_counts, a->peers.size() * 2 * sizeof(int)));
    PICNIX_CUDA(a, cudaMemset(a->d_mig_counts, 0, a->peers.size() * 2 * sizeof(int)));
    for (size_t ip = 0; ip < a->peers.size(); ip++) {
      a->peers[ip].d_psend_count = a->d_mig_counts + 2 * ip;
      a->peers[ip].d_rcount      = a->d_mig_counts + 2 * ip + 1;
    }
  }
  a->nmsg_send_all = (int)sdesc_all.size() / 2;
  a->nmsg_recv_all = (int)rdesc_all.size() / 2;
  if (a->nmsg_send_all > 0 && (status = upload_vector(a, &a->d_send_desc_all, sdesc_all)) != PICNIX_OK)
    return status;
  if (a->nmsg_recv_all > 0 && (status = upload_vector(a, &a->d_recv_desc_all, rdesc_all)) != PICNIX_OK)
    return status;
  for (int mode = 0; mode < 2; mode++) {
    if (a->nmsg_send_all > 0 && (status = upload_vector(a, &a->d_send_off_all[mode], soff_all[mode])) != PICNIX_OK)
      return status;
    if (a->nmsg_recv_all > 0 && (status = upload_vector(a, &a->d_recv_off_all[mode], roff_all[mode])) != PICNIX_OK)
      return status;
  }
  if (!a->slot_peer.empty()) {
    if ((status = upload_vector(a, &a->d_slot_peer, a->slot_peer)) != PICNIX_OK)
      return status;
  }
  // global neighbour ids on the device (destination tags of remote particle records)
  {
    std::vector<int> nbid(a->nbid.begin(), a->nbid.end());
    if ((status = upload_vector(a, &a->d_slot_dst, nbid)) != PICNIX_OK)
      return status;
  }
  return PICNIX_OK;
}

// Particle staging buffers depend on the particle capacity, so they are sized lazily:
// a fraction of the largest local population per peer (records of 64 B).
constexpr int64_t MIG_MIN_BOUND = 16384; // records; smallest message bound of the lagged-count protocol

static int ensure_particle_staging(picnix_arena* a)
{
  if (a->peers.empty() || a->d_psend_ptrs != nullptr)
    return PICNIX_OK;
  int64_t total_cap = 0;
  for (int s = 0; s < a->nseg; s++)
    total_cap += a->seg_cap[s];
  // everything that can leave through one face in a step is far below 1/4 of the population
  int64_t cap = std::max<int64_t>(MIG_MIN_BOUND, total_cap / 4);
  std::vector<double*> ptrs;
  std::vector<int*>    cnts;
  std::vector<int64_t> caps;
  for (auto& p : a->peers) {
    p.pcap_send = cap;
    p.pcap_recv = cap;
    // + 1 record: the lagged-count protocol appends a 64-byte header (the record count)
    PICNIX_CUDA(a, cudaMalloc((void**)&p.d_psend, (cap + 1) * 8 * sizeof(double)));
    PICNIX_CUDA(a, cudaMalloc((void**)&p.d_precv, (cap + 1) * 8 * sizeof(double)));
    ptrs.push_back(p.d_psend);
    cnts.push_back(p.d_psend_count);
    caps.push_back(cap);
  }
  int status;
  if ((status = upload_vector(a, &a->d_psend_ptrs, ptrs)) != PICNIX_OK)
    return status;
  if ((status = upload_vector(a, &a->d_psend_cnts, cnts)) != PICNIX_OK)
    return status;
  if ((status = upload_vector(a, &a->d_psend_caps, caps)) != PICNIX_OK)
    return status;
  PICNIX_CUDA(a, cudaMallocHost((void**)&a->h_mig, a->peers.size() * 2 * sizeof(int)));
  PICNIX_CUDA(a, cudaMallocHost((void**)&a->h_bounds, a->peers.size() * sizeof(int64_t)));
  PICNIX_CUDA(a, cudaEventCreateWithFlags(&a->mig_event, cudaEventDisableTiming));
  return PICNIX_OK;
}

// ---- lagged-count migration protocol -------------------------------------------------------------
// The synchronous protocol reads the number of records per peer back to the host in the middle of
// the step (like MPI_Get_count, nix/chunk.cpp:329-345); that synchronisation exposes the launch
// latency of everything enqueued after it.  Here both sides size the message from the count of the
// PREVIOUS step (sender: what it sent, receiver: what it found in the header -- the same number),
// bound = max(16384, 2 x previous), and the actual count travels in a 64-byte header behind the
// records.  Counts come back to the host one step late, from pinned memory, without a stall.  A step
// that exceeds its bound raises the send-overflow flag (PICNIX_ERR_OVERFLOW at the next
// synchronize) instead of losing particles silently.
__global__ void migration_header_kernel(double** psend, int** counts, const int64_t* bounds, int npeer)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npeer)
    psend[i][bounds[i] * 8] = (double)*counts[i];
}

// the received messages of all peers in one launch: blockIdx.y = peer
constexpr int MAX_PEERS = 32;
struct PeerRecv {
  const double* recv[MAX_PEERS];
  int64_t       bound[MAX_PEERS];
};

__global__ void __launch_bounds__(HALO_THREADS)
unpack_particle_all_kernel(Geom g, DevPtrs d, PeerRecv pr, int* __restrict__ counts, int chunk_begin)
{
  const int     peer  = blockIdx.y;
  const double* recv  = pr.recv[peer];
  const int64_t bound = pr.bound[peer];
  const int     count = (int)recv[bound * 8];
  const int     nrec  = count < bound ? count : (int)bound;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    counts[2 * peer + 1] = count;
    if (count > bound)
      atomicExch(d.errflag + 1, 1);
  }
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrec; r += gridDim.x * blockDim.x) {
    const double* in = recv + (int64_t)r * 8;
    double        p[NC];
#pragma unroll
    for (int k = 0; k < NC; k++)
      p[k] = in[k];
    double tagbits = in[7];
    int2   tag     = *reinterpret_cast<int2*>(&tagbits);
    int    chunk   = tag.x - chunk_begin;
    if (chunk < 0 || chunk >= g.nchunk || tag.y < 0 || tag.y >= g.Ns) {
      atomicExch(d.errflag + 2, 1);
      continue;
    }
    append_particle(g, d, chunk, tag.y, p);
  }
}

// The bound of a message is a function of the previous count ONLY (both sides must compute the same
// number, and their buffer capacities differ): a buffer that is too small for it is grown.
static int ensure_peer_capacity(picnix_arena* a, size_t i, int64_t need)
{
  PeerPlan& p = a->peers[i];
  if (need <= p.pcap_send)
    return PICNIX_OK;
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  cudaFree(p.d_psend);
  cudaFree(p.d_precv);
  PICNIX_CUDA(a, cudaMalloc((void**)&p.d_psend, (need + 1) * 8 * sizeof(double)));
  PICNIX_CUDA(a, cudaMalloc((void**)&p.d_precv, (need + 1) * 8 * sizeof(double)));
  p.pcap_send = p.pcap_recv = need;
  PICNIX_CUDA(a, cudaMemcpy(a->d_psend_ptrs + i, &p.d_psend, sizeof(double*), cudaMemcpyHostToDevice));
  return PICNIX_OK;
}

static bool migration_history_ready(const picnix_arena* a)
{
  for (const auto& p : a->peers)
    if (p.last_sent < 0 || p.last_recv < 0)
      return false;
  return !a->peers.empty();
}

// spilled records with a remote destination (tag.x = ~message slot) -> the peer's send buffer, first in
// line; whatever does not fit (or waits for a local segment) stays on the list
__global__ void __launch_bounds__(HALO_THREADS)
resend_spill_kernel(DevPtrs d, MigrateTables tab, int n, double* __restrict__ keep, int* __restrict__ nkeep)
{
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const double* in      = d.spill_rec + (int64_t)r * 8;
    double        tagbits = in[7];
    const int2    tag     = *reinterpret_cast<int2*>(&tagbits);
    bool          sent    = false;
    if (tag.x < 0) {
      const int peer = tab.slot_peer[~tag.x];
      const int rec  = atomicAdd(tab.psend_cnt[peer], 1);
      if (rec < tab.psend_cap[peer]) {
        double* out = tab.psend[peer] + (int64_t)rec * 8;
#pragma unroll
        for (int k = 0; k < NC; k++)
          out[k] = in[k];
        int2 dst = make_int2(tag.y & 0xffffff, tag.y >> 24); // global chunk id, species
        out[7]   = *reinterpret_cast<double*>(&dst);
        sent     = true;
      } else {
        atomicSub(tab.psend_cnt[peer], 1);
      }
    }
    if (!sent) {
      const int k   = atomicAdd(nkeep, 1);
      double*   out = keep + (int64_t)k * 8;
#pragma unroll
      for (int c = 0; c < 8; c++)
        out[c] = in[c];
    }
  }
}

static int resend_spilled(picnix_arena* a, const MigrateTables& tab)
{
  // the statistics of the previous step say whether anything is waiting (one step late is enough: a
  // record spilled in step n is looked at in step n + 1 by resolve_growth and leaves in step n + 2 at
  // the latest)
  if (a->stat_pending) {
    PICNIX_CUDA(a, cudaEventSynchronize(a->stat_event));
    a->stat_minfree = a->h_stat[0];
    a->stat_maxtail = a->h_stat[1];
    a->stat_spilled = a->h_stat[2];
    a->stat_pending = false;
    a->stat_known   = true;
  }
  if (!a->stat_known || a->stat_spilled <= 0)
    return PICNIX_OK;
  int n = 0;
  PICNIX_CUDA(a, cudaMemcpyAsync(&n, a->d.spill_count, sizeof(int), cudaMemcpyDeviceToHost, a->stream));
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  n = std::min(n, a->d.spill_cap);
  if (n <= 0)
    return PICNIX_OK;
  double* keep  = nullptr;
  int*    nkeep = nullptr;
  PICNIX_CUDA(a, cudaMalloc((void**)&keep, (size_t)n * 8 * sizeof(double)));
  PICNIX_CUDA(a, cudaMalloc((void**)&nkeep, sizeof(int)));
  PICNIX_CUDA(a, cudaMemsetAsync(nkeep, 0, sizeof(int), a->stream));
  resend_spill_kernel<<<std::min(148 * 2, (n + HALO_THREADS - 1) / HALO_THREADS), HALO_THREADS, 0, a->stream>>>(
      a->d, tab, n, keep, nkeep);
  a->kernel_launches++;
  PICNIX_CUDA(a, cudaMemcpyAsync(a->d.spill_rec, keep, (size_t)n * 8 * sizeof(double), cudaMemcpyDeviceToDevice,
                                 a->stream));
  PICNIX_CUDA(a, cudaMemcpyAsync(a->d.spill_count, nkeep, sizeof(int), cudaMemcpyDeviceToDevice, a->stream));
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  cudaFree(keep);
  cudaFree(nkeep);
  return PICNIX_OK;
}

int launch_halo_begin(picnix_arena* a, int mode)
{
  const Geom& g      = a->g;
  int64_t     ncell  = (int64_t)g.nchunk * g.Ng;
  int         blocks = (int)((ncell + HALO_THREADS - 1) / HALO_THREADS);

  switch (mode) {
  case PICNIX_BOUNDARY_EMF: {
    // remote: pack interior margins first (they are not modified by the local gather); the messages of
    // all peers in one launch
    if (a->nmsg_send_all > 0) {
      remote_halo_kernel<6, false, true><<<a->nmsg_send_all, HALO_THREADS, 0, a->stream>>>(
          g, a->d.uf, a->d_send_desc_all, a->d_send_off_all[0], a->d_send_all[0], 0);
      a->kernel_launches++;
    }
    field_halo_local_kernel<<<blocks, HALO_THREADS, 0, a->stream>>>(g, a->d);
    a->kernel_launches++;
    break;
  }
  case PICNIX_BOUNDARY_CUR: {
    // remote: pack ghost regions (read-only for the local gather as well)
    if (a->nmsg_send_all > 0) {
      remote_halo_kernel<4, true, true><<<a->nmsg_send_all, HALO_THREADS, 0, a->stream>>>(
          g, a->d.uj, a->d_send_desc_all, a->d_send_off_all[1], a->d_send_all[1], 0);
      a->kernel_launches++;
    }
    current_halo_local_kernel<<<blocks, HALO_THREADS, 0, a->stream>>>(g, a->d);
    a->kernel_launches++;
    break;
  }
  case PICNIX_BOUNDARY_MOM: {
    // XtensorHaloMoment3D (nix/xtensor_halo3d.hpp:134-185): ghost -> neighbour, added into the
    // interior margin, like the current but with Ns * 14 components per cell
    int status = ensure_moment_array(a);
    if (status != PICNIX_OK)
      return status;
    for (auto& p : a->peers) {
      if (p.d_send[2] == nullptr) {
        PICNIX_CUDA(a, cudaMalloc((void**)&p.d_send[2], std::max<int64_t>(p.send_elems[2], 1) * sizeof(double)));
        PICNIX_CUDA(a, cudaMalloc((void**)&p.d_recv[2], std::max<int64_t>(p.recv_elems[2], 1) * sizeof(double)));
      }
      int nmsg = (int)p.send_chunk.size();
      if (nmsg > 0) {
        remote_halo_kernel<0, true, true><<<nmsg, HALO_THREADS, 0, a->stream>>>(
            g, a->d.um, p.d_send_desc, p.d_send_off[2], p.d_send[2], g.Ns * 14);
        a->kernel_launches++;
      }
    }
    status = launch_moment_halo_local(a);
    if (status != PICNIX_OK)
      return status;
    break;
  }
  case PICNIX_BOUNDARY_PARTICLE: {
    if (!a->particles_allocated)
      return fail(a, PICNIX_ERR_INVALID, "no particles allocated");
    int status = ensure_particle_staging(a);
    if (status != PICNIX_OK)
      return status;
    if ((status = materialize_sort(a)) != PICNIX_OK)
      return status;
    if (!a->peers.empty())
      PICNIX_CUDA(a, cudaMemsetAsync(a->d_mig_counts, 0, a->peers.size() * 2 * sizeof(int), a->stream));
    // lagged-count protocol: bounds from the counts of the previous step (already on the host)
    if (a->mig_pending) {
      PICNIX_CUDA(a, cudaEventSynchronize(a->mig_event)); // recorded a step ago: no stall in steady state
      for (size_t i = 0; i < a->peers.size(); i++) {
        a->peers[i].last_sent = a->h_mig[2 * i + 0];
        a->peers[i].last_recv = a->h_mig[2 * i + 1];
      }
      a->mig_pending = false;
    }
    a->mig_async_step = a->async_migration && migration_history_ready(a);
    for (size_t i = 0; i < a->peers.size(); i++) {
      PeerPlan& p = a->peers[i];
      if (a->mig_async_step) {
        p.send_bound = std::max<int64_t>(MIG_MIN_BOUND, 2 * p.last_sent);
        p.recv_bound = std::max<int64_t>(MIG_MIN_BOUND, 2 * p.last_recv);
        int gstatus  = ensure_peer_capacity(a, i, std::max(p.send_bound, p.recv_bound));
        if (gstatus != PICNIX_OK)
          return gstatus;
        a->h_bounds[i] = p.send_bound;
      } else {
        a->h_bounds[i] = p.pcap_send;
      }
    }
    if (!a->peers.empty())
      PICNIX_CUDA(a, cudaMemcpyAsync(a->d_psend_caps, a->h_bounds, a->peers.size() * sizeof(int64_t),
                                     cudaMemcpyHostToDevice, a->stream));
    if (!a->peers.empty()) {
      // records that did not fit a peer's message in an earlier step leave first
      MigrateTables tab{a->d_slot_peer, a->d_psend_ptrs, a->d_psend_cnts, a->d_psend_caps, a->d_slot_dst};
      if ((status = resend_spilled(a, tab)) != PICNIX_OK)
        return status;
    }
    int maxcap = 0;
    for (int s = 0; s < a->nseg; s++)
      maxcap = std::max(maxcap, a->seg_cap[s]);
    if (maxcap > 0) {
      MigrateTables tab{a->d_slot_peer, a->d_psend_ptrs, a->d_psend_cnts, a->d_psend_caps,
                        a->d_slot_dst};
      const int     use_list = a->leave_list_valid ? 1 : 0;
      const int64_t total    = (int64_t)a->nseg * maxcap;
      if (use_list) {
        migrate_list_kernel<<<148 * 4, HALO_THREADS, 0, a->stream>>>(g, a->d, tab);
        a->kernel_launches++;
      }
      // grid-stride: a resident grid when it is only the overflow fallback of the list
      const int64_t want   = (total + HALO_THREADS - 1) / HALO_THREADS;
      const int     blocks = (int)std::min<int64_t>(want, use_list ? 148 * 8 : 148 * 64);
      migrate_kernel<<<blocks, HALO_THREADS, 0, a->stream>>>(g, a->d, tab, maxcap, use_list);
      a->kernel_launches++;
      a->leave_list_valid = false; // appended migrants are not on the list
    }
    if (a->mig_async_step) {
      // header behind the records, count back to the host for the NEXT step; no synchronisation
      const int npeer = (int)a->peers.size();
      migration_header_kernel<<<1, 32, 0, a->stream>>>(a->d_psend_ptrs, a->d_psend_cnts, a->d_psend_caps, npeer);
      a->kernel_launches++;
      for (int i = 0; i < npeer; i++) {
        PeerPlan& p = a->peers[i];
        // (the counts travel to the host in one copy at the end of the exchange)
        p.psend_bytes = (p.send_bound + 1) * 8 * (int64_t)sizeof(double);
        p.precv_bytes = (p.recv_bound + 1) * 8 * (int64_t)sizeof(double);
      }
    } else if (!a->peers.empty()) {
      // exact send sizes must be known to the host before the transfer (like MPI_Get_count): all
      // peers' counters come back with ONE synchronisation
      std::vector<int> counts(a->peers.size(), 0);
      for (size_t i = 0; i < a->peers.size(); i++)
        PICNIX_CUDA(a, cudaMemcpyAsync(&counts[i], a->peers[i].d_psend_count, sizeof(int),
                                       cudaMemcpyDeviceToHost, a->stream));
      PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
      for (size_t i = 0; i < a->peers.size(); i++) {
        PeerPlan& p   = a->peers[i];
        int64_t count = std::min<int64_t>(counts[i], p.pcap_send);
        p.psend_bytes = count * 8 * (int64_t)sizeof(double);
        p.precv_bytes = 0;
        p.last_sent   = count;
      }
    }
    break;
  }
  default:
    return fail(a, PICNIX_ERR_INVALID, "No such boundary mode exists!");
  }
  return check_cuda(a, cudaGetLastError(), "boundary_begin");
}

int launch_halo_end(picnix_arena* a, int mode)
{
  const Geom& g = a->g;
  switch (mode) {
  case PICNIX_BOUNDARY_EMF:
    if (a->nmsg_recv_all > 0) {
      remote_halo_kernel<6, false, false><<<a->nmsg_recv_all, HALO_THREADS, 0, a->stream>>>(
          g, a->d.uf, a->d_recv_desc_all, a->d_recv_off_all[0], a->d_recv_all[0], 0);
      a->kernel_launches++;
    }
    break;
  case PICNIX_BOUNDARY_CUR:
    if (a->nmsg_recv_all > 0) {
      remote_halo_kernel<4, true, false><<<a->nmsg_recv_all, HALO_THREADS, 0, a->stream>>>(
          g, a->d.uj, a->d_recv_desc_all, a->d_recv_off_all[1], a->d_recv_all[1], 0);
      a->kernel_launches++;
    }
    break;
  case PICNIX_BOUNDARY_MOM:
    for (auto& p : a->peers) {
      int nmsg = (int)p.recv_chunk.size();
      if (nmsg > 0 && p.d_recv[2] != nullptr) {
        remote_halo_kernel<0, true, false><<<nmsg, HALO_THREADS, 0, a->stream>>>(
            g, a->d.um, p.d_recv_desc, p.d_recv_off[2], p.d_recv[2], g.Ns * 14);
        a->kernel_launches++;
      }
    }
    break;
  case PICNIX_BOUNDARY_PARTICLE: {
    if (a->mig_async_step) {
      const int npeer = (int)a->peers.size();
      if (npeer > MAX_PEERS)
        return fail(a, PICNIX_ERR_INVALID, "more than 32 peer ranks");
      if (npeer > 0) {
        PeerRecv pr;
        for (int i = 0; i < npeer; i++) {
          pr.recv[i]  = a->peers[i].d_precv;
          pr.bound[i] = a->peers[i].recv_bound;
        }
        unpack_particle_all_kernel<<<dim3(148, npeer), HALO_THREADS, 0, a->stream>>>(g, a->d, pr, a->d_mig_counts,
                                                                                    a->chunk_begin);
        a->kernel_launches++;
        // sent and received counts of all peers for the next step's bounds
        PICNIX_CUDA(a, cudaMemcpyAsync(a->h_mig, a->d_mig_counts, npeer * 2 * sizeof(int), cudaMemcpyDeviceToHost,
                                       a->stream));
      }
      PICNIX_CUDA(a, cudaEventRecord(a->mig_event, a->stream));
      a->mig_pending = true;
    } else {
      for (auto& p : a->peers) {
        int nrec = (int)(p.precv_bytes / (8 * sizeof(double)));
        if (nrec > 0) {
          unpack_particle_kernel<<<(nrec + HALO_THREADS - 1) / HALO_THREADS, HALO_THREADS, 0,
                                   a->stream>>>(g, a->d, p.d_precv, nrec, a->chunk_begin);
          a->kernel_launches++;
        }
        p.last_recv = nrec;
      }
    }
    // pre_unpack would have resized the particle arrays (nix/xtensor_halo3d.hpp:406-418): grow the segments
    // that are (nearly) full and append the migrants that were waiting for room
    int status = resolve_growth(a);
    if (status != PICNIX_OK)
      return status;
    // post_unpack ends with sort() for every species (nix/xtensor_halo3d.hpp:495-497)
    status = launch_sort(a, 0, -1);
    if (status != PICNIX_OK)
      return status;
    break;
  }
  default:
    return fail(a, PICNIX_ERR_INVALID, "No such boundary mode exists!");
  }
  {
    // PicChunk::set_boundary_unpack ends with the physical boundary hook (pic/pic_chunk.cpp:360-361)
    int status = launch_boundary_field(a, mode);
    if (status != PICNIX_OK)
      return status;
  }
  return check_cuda(a, cudaGetLastError(), "boundary_end");
}

// PicChunk::inject_particle hook (called by set_boundary_pack(BoundaryParticle), pic/pic_chunk.cpp:305-310):
// particles the host generated (example/shock/main.cpp:436-535 draws them from the host's generators) are
// appended behind the active particles of (chunk, species), keyed and counted, before the exchange's sort
__global__ void __launch_bounds__(HALO_THREADS)
inject_kernel(Geom g, DevPtrs d, const double* __restrict__ aos, int n, int chunk, int is)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n)
    return;
  double p[NC];
#pragma unroll
  for (int k = 0; k < NC; k++)
    p[k] = aos[(int64_t)r * NC + k];
  append_particle(g, d, chunk, is, p);
}

int inject_particles(picnix_arena* a, int ichunk, int is, const double* aos, int n)
{
  if (n <= 0)
    return PICNIX_OK;
  double* dbuf = nullptr;
  PICNIX_CUDA(a, cudaMalloc((void**)&dbuf, (size_t)n * NC * sizeof(double)));
  PICNIX_CUDA(a, cudaMemcpyAsync(dbuf, aos, (size_t)n * NC * sizeof(double), cudaMemcpyHostToDevice, a->stream));
  inject_kernel<<<(n + HALO_THREADS - 1) / HALO_THREADS, HALO_THREADS, 0, a->stream>>>(a->g, a->d, dbuf, n, ichunk, is);
  a->kernel_launches++;
  PICNIX_CUDA(a, cudaStreamSynchronize(a->stream));
  cudaFree(dbuf);
  a->leave_list_valid = false;
  return check_cuda(a, cudaGetLastError(), "inject_particles");
}

} // namespace picnix

using namespace picnix;

extern "C" {

int picnix_cuda_inject_particles(picnix_arena_t* a, int32_t ichunk, int32_t is, const double* xu_aos, int32_t n)
{
  if (a == nullptr || ichunk < 0 || ichunk >= a->g.nchunk || is < 0 || is >= a->g.Ns || n < 0 ||
      (xu_aos == nullptr && n > 0) || !a->particles_allocated)
    return PICNIX_ERR_INVALID;
  int status = materialize_sort(a);
  if (status != PICNIX_OK)
    return status;
  return inject_particles(a, ichunk, is, xu_aos, n);
}

int picnix_cuda_get_peers(const picnix_arena_t* a, int32_t* npeer, int32_t* peer_rank)
{
  if (a == nullptr || npeer == nullptr)
    return PICNIX_ERR_INVALID;
  *npeer = (int32_t)a->peers.size();
  if (peer_rank != nullptr) {
    for (size_t i = 0; i < a->peers.size(); i++)
      peer_rank[i] = a->peers[i].rank;
  }
  return PICNIX_OK;
}

int picnix_cuda_get_comm_buffer(picnix_arena_t* a, int32_t mode, int32_t peer_index,
                                void** send_ptr, int64_t* send_bytes, void** recv_ptr,
                                int64_t* recv_bytes)
{
  if (a == nullptr || peer_index < 0 || peer_index >= (int)a->peers.size())
    return PICNIX_ERR_INVALID;
  PeerPlan& p = a->peers[peer_index];
  if (mode == PICNIX_BOUNDARY_EMF || mode == PICNIX_BOUNDARY_CUR || mode == PICNIX_BOUNDARY_MOM) {
    *send_ptr   = p.d_send[mode];
    *recv_ptr   = p.d_recv[mode];
    *send_bytes = p.send_elems[mode] * (int64_t)sizeof(double);
    *recv_bytes = p.recv_elems[mode] * (int64_t)sizeof(double);
    return PICNIX_OK;
  }
  if (mode == PICNIX_BOUNDARY_PARTICLE) {
    int status = ensure_particle_staging(a);
    if (status != PICNIX_OK)
      return status;
    *send_ptr   = p.d_psend;
    *recv_ptr   = p.d_precv;
    *send_bytes = p.psend_bytes;
    *recv_bytes = p.precv_bytes;
    return PICNIX_OK;
  }
  return fail(a, PICNIX_ERR_INVALID, "No such boundary mode exists!");
}

int picnix_cuda_set_recv_bytes(picnix_arena_t* a, int32_t mode, int32_t peer_index,
                               int64_t recv_bytes)
{
  if (a == nullptr || peer_index < 0 || peer_index >= (int)a->peers.size() ||
      mode != PICNIX_BOUNDARY_PARTICLE)
    return PICNIX_ERR_INVALID;
  PeerPlan& p = a->peers[peer_index];
  if (recv_bytes < 0 || recv_bytes > (p.pcap_recv + 1) * 8 * (int64_t)sizeof(double))
    return fail(a, PICNIX_ERR_OVERFLOW, "particle receive buffer too small");
  p.precv_bytes = recv_bytes;
  return PICNIX_OK;
}

} // extern "C"
