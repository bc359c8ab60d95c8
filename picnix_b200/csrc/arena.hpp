// -*- C++ -*-
// Internal definitions shared by the CUDA translation units of libpicnix_b200.so.
//
// Nothing here is part of the C ABI (include/picnix_b200.h); it is the device-side data model:
// one arena per GPU holding every chunk of the rank in batched arrays, so that each phase of
// PicApplication::push_openmp (pic/pic_application.cpp:219-292) is ONE kernel launch over all
// chunks instead of one host call per chunk.
#ifndef PICNIX_B200_ARENA_HPP
#define PICNIX_B200_ARENA_HPP

#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/picnix_b200.h"

namespace picnix
{

constexpr int NC         = 7;   // components per particle, nix/particle.hpp:18
constexpr int ALLOC_UNIT = 128; // nix/particle.hpp:19
constexpr int NBSIZE     = 27;  // nix/chunk.hpp:18

// neighbour table codes (non-negative = local chunk index)
constexpr int NB_NONE        = -1; // MPI_PROC_NULL: no neighbour (non-periodic face)
constexpr int NB_REMOTE_BASE = -2; // -(2 + message slot): neighbour lives on another rank

// Geometry common to all chunks of an arena; passed BY VALUE to every kernel.
// Index order is (z, y, x) everywhere, as in the reference.
struct Geom {
  int    nchunk;     // local chunks
  int    Ns;         // species
  int    dims[3];    // cells per chunk
  int    nb;         // boundary margin = (order+3)/2, pic/pic_chunk.cpp:195
  int    M[3];       // padded extents dims + 2*nb (1 + 2*nb for ignorable dims)
  int    Lb[3];      // first interior index
  int    Ub[3];      // last interior index
  int    has_dim[3]; // 0 for ignorable dimensions
  int    dimension;  // 1, 2 or 3
  int    order;      // shape order
  int    is_odd;     // order % 2
  int    Ng;         // Mz*My*Mx: number of bins of the counting sort, nix/particle.hpp:100-106
  int    fsy;        // flatindex stride in y = Ubx-Lbx+2, nix/xtensor_particle.hpp:231-238
  int    fsz;        // flatindex stride in z = fsy*(Uby-Lby+2)
  double cc;
  double del[3];     // dz, dy, dx
  double glim[3][2]; // global domain [z|y|x][min|max], nix/chunk.cpp:210-237
  double theta;      // Friedman filter parameter
  // physical boundary conditions on non-periodic faces (boundary.cu): kind and values per [z|y|x][lower|upper]
  int    bc_kind[3][2];
  int    any_particle_bc; // some face reflects particles
  double bc_val[3][2][6]; // PICNIX_BC_INFLOW: Ex, Ey, Ez, Bx, By, Bz imposed in the margin
};

// Per-arena device pointers; passed BY VALUE to kernels.
struct DevPtrs {
  double*  uf;       // [nchunk][Mz][My][Mx][6]
  double*  uj;       // [nchunk][Mz][My][Mx][4]
  double*  ff;       // [nchunk][Mz][My][Mx][3][3]  (time level, E component)
  double*  um;       // [nchunk][Mz][My][Mx][Ns][14] velocity moments; allocated by the first deposit_moment
  double*  clim;     // [nchunk][3][2] chunk limits  [z|y|x][min|max] (actual, not +-DBL_MAX)
  int*     nbr;      // [nchunk][27] neighbour codes
  double*  xu;       // [7][pcap]  SoA particle buffer "xu"
  double*  xv;       // [7][pcap]  SoA particle buffer "xv"
  int*     gindex;   // [pcap] cell key of xu
  int*     perm;     // [pcap] lazy sort: sorted slot -> slot the particle still sits in (per segment)
  int64_t  pcap;     // component stride of xu/xv
  int64_t* seg_off;  // [nseg] first slot of segment (chunk*Ns + species)
  int*     seg_cap;  // [nseg] capacity
  int*     np;       // [nseg] active particles
  int*     ntail;    // [nseg] migrants appended behind np during the current step
  int*     pindex;   // [nseg][Ng+1]
  int*     pcount;   // [nseg][Ng+1] histogram / scatter cursor
  double*  qm;       // [Ns][2] charge, mass
  int*     errflag;  // [4] device-side error flags (0: segment overflow, 1: send overflow,
                     //     2: bad migration tag, 3: far-mover list overflow)
  int*     far_count; // [1] particles that moved more than one cell in the row kernel
  double*  far_rec;   // [far_cap][8] x0,y0,z0,x1,y1,z1,q,chunk: deposited by a follow-up kernel
  int      far_cap;
  // particles that left their chunk in the fused push (key == Ng), as (segment << 40 | slot):
  // the migration visits this list instead of scanning every key again
  int*     leave_count; // [1]
  int64_t* leave_idx;   // [leave_cap]
  int      leave_cap;
  // migrants that found their destination full (a segment, or a peer's message bound): kept here
  // instead of being dropped and re-appended after the segments have grown (grow.cu)
  int*     spill_count; // [1]
  double*  spill_rec;   // [spill_cap][8] 7 components + tag (int2: destination segment | ~message slot, species)
  int      spill_cap;
  int*     seg_stat;    // [4] min over segments of (capacity - np), max ntail, spill count, -- ; written by the sort
};

struct PeerPlan {
  int                  rank;      // peer rank
  // fixed-size modes (Emf, Cur): list of (local chunk, direction) messages in a canonical order
  std::vector<int>     send_chunk, send_dir; // what we send
  std::vector<int>     recv_chunk, recv_dir; // what we receive (same canonical order on the peer)
  int*                 d_send_desc = nullptr; // device copy [nmsg][2]
  int*                 d_recv_desc = nullptr;
  int64_t              send_elems[3] = {0, 0, 0}; // doubles per mode (Emf, Cur, Mom)
  int64_t              recv_elems[3] = {0, 0, 0};
  std::vector<int64_t> send_msg_off[3], recv_msg_off[3]; // element offset of each message
  int64_t*             d_send_off[3] = {nullptr, nullptr, nullptr};
  int64_t*             d_recv_off[3] = {nullptr, nullptr, nullptr};
  double*              d_send[3] = {nullptr, nullptr, nullptr}; // Mom buffers are allocated on first use
  double*              d_recv[3] = {nullptr, nullptr, nullptr};
  // particle mode: fixed-capacity staging, records of 8 doubles (7 comps + destination code)
  double*              d_psend = nullptr;
  double*              d_precv = nullptr;
  int*                 d_psend_count = nullptr; // device counter
  int64_t              pcap_send = 0, pcap_recv = 0;
  int64_t              psend_bytes = 0, precv_bytes = 0;
  // lagged-count migration protocol (option "async_migration"): records sent / received in the
  // previous step (-1: unknown) and the message bounds both sides derive from them
  int64_t              last_sent = -1, last_recv = -1;
  int64_t              send_bound = 0, recv_bound = 0;
  int*                 d_rcount = nullptr; // device copy of the count found in the received header
};

} // namespace picnix

namespace picnix
{
struct HostIO; // hostio.cu: streams, slabs and events of the pipelined host-buffer step
}

// The opaque handle of the C ABI.
struct picnix_arena {
  picnix_config_t        cfg;
  picnix::Geom           g;
  picnix::DevPtrs        d;
  cudaStream_t           stream      = nullptr;
  bool                   own_stream  = false;
  int                    nseg        = 0;
  int                    chunk_begin = 0; // first global chunk id
  int                    nchunk_global = 0;
  std::vector<int32_t>   chunkid;     // [Cz][Cy][Cx]
  std::vector<int32_t>   coord;       // [nchunk_global][3] (x,y,z)
  std::vector<int32_t>   boundary;    // rank boundary
  std::vector<int32_t>   nbid;        // [nchunk][27] global neighbour ids
  std::vector<int32_t>   nbrank;      // [nchunk][27]
  std::vector<int32_t>   nbr_code;    // [nchunk][27] host copy of DevPtrs::nbr
  std::vector<int64_t>   seg_off;     // host copies
  std::vector<int32_t>   seg_cap;
  std::vector<picnix::PeerPlan> peers;
  std::vector<int32_t>   slot_peer;   // remote message slot -> peer index
  std::vector<int32_t>   slot_dst;    // remote message slot -> destination chunk's local index on the peer
  int*                   d_slot_peer = nullptr;
  int*                   d_slot_dst  = nullptr;
  double**               d_psend_ptrs = nullptr; // [npeer] device table of particle send buffers
  int**                  d_psend_cnts = nullptr; // [npeer]
  int64_t*               d_psend_caps = nullptr; // [npeer]
  void*                  d_scan_tmp  = nullptr;
  double*                d_reduce    = nullptr; // [nchunk][4] reduction scratch
  double*                h_stage     = nullptr; // pinned staging for AoS<->SoA transfers
  double*                d_stage     = nullptr;
  int64_t                stage_elems = 0;
  picnix::HostIO*        hostio      = nullptr;
  bool                   particles_allocated = false;
  bool                   pindex_valid = false;  // pindex matches the particle order (after a sort)
  bool                   leave_list_valid = false; // DevPtrs::leave_idx describes the current keys
  bool                   lazy_sort    = true;   // option "lazy_sort": allow index-only sorts (see sort.cu)
  bool                   async_migration = false; // option "async_migration": no host sync in the particle exchange
  bool                   mig_async_step  = false; // the exchange in flight uses the lagged-count protocol
  int*                   h_mig      = nullptr;    // pinned [npeer][2]: sent / received counts of the last step
  int64_t*               h_bounds   = nullptr;    // pinned [npeer]: per-step send bounds (device caps)
  // fixed-size modes (Emf, Cur): the buffers of all peers are carved from one allocation per mode and
  // direction, and the messages of all peers are listed in one table, so that a phase is ONE pack or
  // unpack launch whatever the number of peers
  double*                d_send_all[2] = {nullptr, nullptr};
  double*                d_recv_all[2] = {nullptr, nullptr};
  int*                   d_send_desc_all = nullptr;
  int*                   d_recv_desc_all = nullptr;
  int64_t*               d_send_off_all[2] = {nullptr, nullptr};
  int64_t*               d_recv_off_all[2] = {nullptr, nullptr};
  int                    nmsg_send_all = 0, nmsg_recv_all = 0;
  int*                   d_mig_counts  = nullptr; // [npeer][2]: records sent / found in the received header
  cudaEvent_t            mig_event  = nullptr;
  bool                   mig_pending = false;     // h_mig will hold the counts of the last step after mig_event
  bool                   perm_pending = false;  // xu is NOT yet in pindex order: DevPtrs::perm holds the order
  bool                   any_bc        = false; // some face has a physical boundary condition (boundary.cu)
  bool                   force_generic = false; // testing: bypass the tiled kernels
  bool                   deposit_mma   = false; // row kernel variant: deposit through the FP64 MMA unit
  int                    row_version   = 2;     // option "row_kernel": 2 = rowpush.cu, 1 = round-1 kernel (rowfused.cu)
  // growth of particle segments (grow.cu): statistics of the previous step arrive one step late
  int*                   h_stat       = nullptr; // pinned [4] copy of DevPtrs::seg_stat
  cudaEvent_t            stat_event   = nullptr;
  bool                   stat_pending = false;
  bool                   stat_known   = false;
  int                    stat_minfree = 0, stat_maxtail = 0, stat_spilled = 0;
  bool                   check_growth_always = false; // option "check_growth": host check before every sort
  int64_t                segment_regrows = 0;   // how often the particle arrays were re-laid out
  int64_t                late_particles  = 0;   // migrants delivered one step late (spilled in the unchecked path)
  int64_t                kernel_launches = 0;
  int64_t                particle_pushes = 0;
  int64_t                np_total_hint   = 0; // sum of np at last host-visible count
  std::string            error;
};

namespace picnix
{

// error helpers --------------------------------------------------------------------------------
int  fail(picnix_arena* a, int code, const std::string& msg);
int  check_cuda(picnix_arena* a, cudaError_t err, const char* what);
#define PICNIX_CUDA(a, call)                                                                      \
  do {                                                                                             \
    int status_ = picnix::check_cuda((a), (call), #call);                                          \
    if (status_ != PICNIX_OK)                                                                      \
      return status_;                                                                              \
  } while (0)

inline void resolve_range(const picnix_arena* a, int& c0, int& cn)
{
  if (cn < 0) {
    c0 = 0;
    cn = a->g.nchunk;
  }
}

// host decomposition (decomp.cpp) -----------------------------------------------------------------
void sfc_build(int Cz, int Cy, int Cx, std::vector<int32_t>& chunkid, std::vector<int32_t>& coord);
bool assign_binarysearch(const std::vector<double>& load, std::vector<int32_t>& boundary);
bool assign_smilei(const std::vector<double>& load, std::vector<int32_t>& boundary);
std::vector<int32_t> assign_initial(const std::vector<double>& load, int nrank);

// phase launchers (one per .cu) ---------------------------------------------------------------------
int launch_init_friedman(picnix_arena* a, int c0, int cn);
int launch_push_bfd(picnix_arena* a, int c0, int cn, double delt);
int launch_push_efd(picnix_arena* a, int c0, int cn, double delt);
int launch_diverror(picnix_arena* a, double* efd, double* bfd);
int launch_field_energy(picnix_arena* a, double* efd, double* bfd);

int launch_push_velocity(picnix_arena* a, int c0, int cn, double delt);
int launch_push_position(picnix_arena* a, int c0, int cn, double delt);
int launch_deposit_current(picnix_arena* a, int c0, int cn, double delt);
int launch_push_deposit_fused(picnix_arena* a, int c0, int cn, double delt);
int launch_deposit_rows(picnix_arena* a, int c0, int cn, double delt);
bool row_kernel_applies(const picnix_arena* a);
bool row_geometry_applies(const picnix_arena* a);

int ensure_moment_array(picnix_arena* a);
int launch_deposit_moment(picnix_arena* a);
int launch_moment_halo_local(picnix_arena* a);
int launch_particle_energy(picnix_arena* a, double* particle);

int launch_count(picnix_arena* a, int c0, int cn);
int launch_sort(picnix_arena* a, int c0, int cn);
int materialize_sort(picnix_arena* a); // physically order xu if an index-only sort is pending

int resolve_growth(picnix_arena* a);  // grow.cu: before the sort of the particle exchange
int record_segment_stats(picnix_arena* a);
int launch_boundary_field(picnix_arena* a, int mode); // boundary.cu: PicChunk::set_boundary_field
int launch_halo_begin(picnix_arena* a, int mode);
int launch_halo_end(picnix_arena* a, int mode);

void hostio_destroy(picnix_arena* a);
int  upload_state_pipelined(picnix_arena* a, double* uf, double* uj, double* ff, double* xu,
                            const int32_t* np_in, const int32_t* np_cap, bool with_uj);
int  download_state_pipelined(picnix_arena* a, double* uf, double* uj, double* ff, double* xu,
                              const int32_t* np_cap, int32_t* np_out);
int  step_host_pipelined(picnix_arena* a, double delt, int nstep, double* uf, double* uj, double* ff,
                         double* xu, const int32_t* np_in, const int32_t* np_cap, int32_t* np_out);

int upload_particles(picnix_arena* a, int ichunk, int is, const double* aos, int np);
int download_particles(picnix_arena* a, int ichunk, int is, int which, int n, double* aos);

} // namespace picnix

#endif
