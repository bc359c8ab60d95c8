/*
 * picnix_oracle.h -- plain-C restatement of the PIC-NIX per-timestep hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle ("port"): a scalar CPU restatement of the
 * reference's algorithm (amanotk/pic-nix @ 9c960d5), each function citing the reference file:line
 * it follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * leg may load it; the product library (libpicnix_b200.so) never links, loads or calls it.
 *
 * Parity is PINNED: tests/test_oracle_vs_reference.py checks this restatement against the
 * UNMODIFIED reference compiled here (oracle/_ref, see oracle/Makefile), and tests/golden/ holds
 * vectors generated from that compiled reference (tests/golden/make_golden.py) so the pin also
 * holds where /root/reference and oracle/_ref are absent.
 *
 * Layouts are the reference's: uf[Mz][My][Mx][6], uj[..][4], ff[..][3][6], particles AoS [np][7].
 */
#ifndef PICNIX_ORACLE_H
#define PICNIX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_config {
  int32_t ndims[3];    /* global cells z,y,x                                   */
  int32_t cdims[3];    /* chunks z,y,x                                         */
  int32_t periodic[3]; /* z,y,x                                                */
  int32_t order;       /* 1..4                                                 */
  int32_t pusher;      /* 0 Boris, 1 Vay, 2 Higuera-Cary                       */
  int32_t interp;      /* 0 MC, 1 WT                                           */
  int32_t Ns;
  int32_t nrank, rank; /* chunk ids are split over nrank ranks; this is `rank` */
  int32_t simd_width;  /* stripe width of the counting sort (NIX_SIMD_WIDTH)   */
  int32_t nthread;     /* OpenMP threads over chunks (<=0: all)                */
  double  cc, delx, dely, delz, friedman, buffer_ratio;
} orc_config_t;

typedef struct orc_sim orc_sim_t;

/* decomposition (nix/sfc.cpp, nix/balancer.cpp, nix/chunkmap.cpp) */
int orc_sfc_build(int32_t Cz, int32_t Cy, int32_t Cx, int32_t* chunkid, int32_t* coord);
int orc_assign_initial(const double* load, int32_t nchunk, int32_t nrank, int32_t* boundary);
int orc_assign_rebalance(const double* load, int32_t nchunk, int32_t nrank, int32_t* boundary);

/* boundary: nrank+1 ascending chunk ids or NULL (even split) */
orc_sim_t* orc_create(const orc_config_t* cfg, const int32_t* boundary);
void       orc_destroy(orc_sim_t* s);
int        orc_num_chunks(const orc_sim_t* s);
int        orc_num_threads(const orc_sim_t* s);
int        orc_chunk_id_begin(const orc_sim_t* s);
void       orc_get_shape(const orc_sim_t* s, int32_t* shape5 /* Mz,My,Mx,nb,Ng */);
void       orc_get_neighbors(const orc_sim_t* s, int ic, int32_t* nbid, int32_t* nbrank);

void orc_set_species(orc_sim_t* s, int is, double q, double m);
void orc_set_field(orc_sim_t* s, int ic, int which, const double* in);
void orc_get_field(const orc_sim_t* s, int ic, int which, double* out);
void orc_set_particles(orc_sim_t* s, int ic, int is, const double* xu, int np, int np_alloc);
int  orc_get_np(const orc_sim_t* s, int ic, int is);
void orc_get_particles(const orc_sim_t* s, int ic, int is, int which, int n, double* out);
void orc_get_pindex(const orc_sim_t* s, int ic, int is, int32_t* out);
void orc_get_gindex(const orc_sim_t* s, int ic, int is, int n, int32_t* out);

void orc_init_friedman(orc_sim_t* s);
void orc_push_bfd(orc_sim_t* s, double delt);
void orc_push_efd(orc_sim_t* s, double delt);
void orc_push_velocity(orc_sim_t* s, double delt);
void orc_push_position(orc_sim_t* s, double delt);
void orc_deposit_current(orc_sim_t* s, double delt);
void orc_deposit_moment(orc_sim_t* s);
void orc_sort_particle(orc_sim_t* s);

/* boundary exchange: begin packs the per-peer send buffers, end consumes the receive buffers */
void orc_boundary_begin(orc_sim_t* s, int mode);
void orc_boundary_end(orc_sim_t* s, int mode);
int  orc_get_peers(const orc_sim_t* s, int32_t* peer_rank);
void orc_get_comm_buffer(orc_sim_t* s, int mode, int peer_index, void** send_ptr, int64_t* send_bytes,
                         void** recv_ptr, int64_t* recv_bytes);
void orc_set_recv_bytes(orc_sim_t* s, int mode, int peer_index, int64_t recv_bytes);

/* single-rank conveniences */
void orc_exchange(orc_sim_t* s, int mode);
void orc_step(orc_sim_t* s, double delt, int nstep);

void orc_get_diverror(const orc_sim_t* s, int ic, double* efd, double* bfd);
void orc_get_energy(const orc_sim_t* s, int ic, double* efd, double* bfd, double* particle);

#ifdef __cplusplus
}
#endif
#endif
