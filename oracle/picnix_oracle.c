/*
 * picnix_oracle.c -- plain-C restatement of the PIC-NIX per-timestep hot path (see picnix_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: the parity checker, never the thing measured or shipped.
 * Compiled with -ffp-contract=off, scalar code, reference array layouts.  Every routine follows the
 * reference's SCALAR code path and cites it (paths relative to amanotk/pic-nix @ 9c960d5).
 */
#include "picnix_oracle.h"

#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NC 7          /* nix/particle.hpp:18 */
#define ALLOC_UNIT 128 /* nix/particle.hpp:19 */
#define MAXS 7        /* Order + 3 for Order <= 4 */

enum { MODE_EMF = 0, MODE_CUR = 1, MODE_MOM = 2, MODE_PARTICLE = 3 };
enum { FIELD_UF = 0, FIELD_UJ = 1, FIELD_FF = 2, FIELD_UM = 3 };

typedef struct {
  double  q, m;
  int     np, cap, ntail;
  double *xu, *xv;
  int    *gindex, *pindex, *pcount; /* pcount[(Ng+1)][W] */
} species_t;

typedef struct {
  int        id;
  int        coord[3]; /* z,y,x */
  int        nbid[27], nbrank[27];
  double     lim[3][2]; /* z,y,x : min,max */
  double    *uf, *uj, *ff, *um;
  species_t *sp;
  /* outgoing particles of the current exchange: per direction, per species list of indices */
  int *out_idx[27];
  int *out_cnt[27]; /* [Ns] */
  int  out_cap[27];
} chunk_t;

typedef struct {
  int      rank;
  int      nsend, nrecv;
  int     *send_chunk, *send_dir, *recv_chunk, *recv_dir; /* canonical order */
  int64_t *send_off[3], *recv_off[3];                     /* element offsets per mode (EMF, CUR, MOM) */
  int64_t  send_elems[3], recv_elems[3];
  double  *send[3], *recv[3];
  double  *psend, *precv;
  int64_t  psend_cap, precv_cap, psend_bytes, precv_bytes;
} peer_t;

struct orc_sim {
  orc_config_t cfg;
  int          nb, order, is_odd, dimension, W;
  int          dims[3], M[3], Lb[3], Ub[3], has_dim[3];
  int          Ng, fsy, fsz;
  double       del[3], glim[3][2];
  int          nchunk_global, nchunk, chunk_begin, nthread;
  int32_t     *chunkid, *coord, *boundary;
  chunk_t     *chunks;
  int          npeer;
  peer_t      *peers;
  int         *peer_of_rank;
};

/* ------------------------------------------------------------------------------------------ */
/* decomposition                                                                              */
/* ------------------------------------------------------------------------------------------ */

static int isgn(int v) { return (v > 0) - (v < 0); }

typedef struct {
  int32_t* index;
  int      Ny, Nx, id;
} walk_t;

static void walk_visit(walk_t* w, const int p[3])
{
  w->index[(size_t)(p[2] * w->Ny + p[1]) * w->Nx + p[0]] = w->id++;
}

static void walk_line(walk_t* w, const int p0[3], const int step[3], int n)
{
  int p[3] = {p0[0], p0[1], p0[2]};
  for (int i = 0; i < n; i++) {
    walk_visit(w, p);
    for (int k = 0; k < 3; k++)
      p[k] += step[k];
  }
}

#define V3_SET(d, a)                                                                               \
  do {                                                                                             \
    (d)[0] = (a)[0];                                                                               \
    (d)[1] = (a)[1];                                                                               \
    (d)[2] = (a)[2];                                                                               \
  } while (0)

static int v3len(const int a[3]) { return abs(a[0] + a[1] + a[2]); }

/* generalized Hilbert curve in a plane spanned by a (major) and b (minor): nix/sfc.cpp:200-294 */
static void gilbert2(walk_t* wk, const int p[3], const int a[3], const int b[3])
{
  int w = v3len(a), h = v3len(b);
  int da[3], db[3], a2[3], b2[3], q[3], t1[3], t2[3];
  for (int k = 0; k < 3; k++) {
    da[k] = isgn(a[k]);
    db[k] = isgn(b[k]);
    a2[k] = a[k] / 2;
    b2[k] = b[k] / 2;
  }
  if (h == 1) {
    walk_line(wk, p, da, w);
    return;
  }
  if (w == 1) {
    walk_line(wk, p, db, h);
    return;
  }
  int w2 = v3len(a2), h2 = v3len(b2);
  if (2 * w > 3 * h) {
    if ((w2 % 2) && (w > 2))
      for (int k = 0; k < 3; k++)
        a2[k] += da[k];
    gilbert2(wk, p, a2, b);
    for (int k = 0; k < 3; k++) {
      q[k]  = p[k] + a2[k];
      t1[k] = a[k] - a2[k];
    }
    gilbert2(wk, q, t1, b);
  } else {
    if ((h2 % 2) && (h > 2))
      for (int k = 0; k < 3; k++)
        b2[k] += db[k];
    gilbert2(wk, p, b2, a2);
    for (int k = 0; k < 3; k++) {
      q[k]  = p[k] + b2[k];
      t1[k] = b[k] - b2[k];
    }
    gilbert2(wk, q, a, t1);
    for (int k = 0; k < 3; k++) {
      q[k]  = p[k] + (a[k] - da[k]) + (b2[k] - db[k]);
      t1[k] = -b2[k];
      t2[k] = -(a[k] - a2[k]);
    }
    gilbert2(wk, q, t1, t2);
  }
}

/* three-dimensional generalized Hilbert curve: nix/sfc.cpp:296-508 */
static void gilbert3(walk_t* wk, const int p0[3], const int a[3], const int b[3], const int c[3])
{
  int w = v3len(a), h = v3len(b), d = v3len(c);
  int da[3], db[3], dc[3], a2[3], b2[3], c2[3], a3[3], b3[3], c3[3], p[3], n1[3], n2[3], n3[3];
  for (int k = 0; k < 3; k++) {
    da[k] = isgn(a[k]);
    db[k] = isgn(b[k]);
    dc[k] = isgn(c[k]);
    a2[k] = a[k] / 2;
    b2[k] = b[k] / 2;
    c2[k] = c[k] / 2;
    p[k]  = p0[k];
  }
  if (h == 1 && d == 1) {
    walk_line(wk, p, da, w);
    return;
  }
  if (w == 1 && d == 1) {
    walk_line(wk, p, db, h);
    return;
  }
  if (w == 1 && h == 1) {
    walk_line(wk, p, dc, d);
    return;
  }
  if ((v3len(a2) % 2) && (w > 2))
    for (int k = 0; k < 3; k++)
      a2[k] += da[k];
  if ((v3len(b2) % 2) && (h > 2))
    for (int k = 0; k < 3; k++)
      b2[k] += db[k];
  if ((v3len(c2) % 2) && (d > 2))
    for (int k = 0; k < 3; k++)
      c2[k] += dc[k];
  for (int k = 0; k < 3; k++) {
    a3[k] = a[k] - a2[k];
    b3[k] = b[k] - b2[k];
    c3[k] = c[k] - c2[k];
  }

  if ((2 * w > 3 * h) && (2 * w > 3 * d)) {
    gilbert3(wk, p, a2, b, c);
    for (int k = 0; k < 3; k++)
      p[k] += a2[k];
    gilbert3(wk, p, a3, b, c);
  } else if (3 * h > 4 * d) {
    gilbert3(wk, p, b2, c, a2);
    for (int k = 0; k < 3; k++)
      p[k] += b2[k];
    gilbert3(wk, p, a, b3, c);
    for (int k = 0; k < 3; k++) {
      p[k] += (a[k] - da[k]) - db[k];
      n1[k] = -b2[k];
      n2[k] = -a3[k];
    }
    gilbert3(wk, p, n1, c, n2);
  } else if (3 * d > 4 * h) {
    gilbert3(wk, p, c2, a2, b);
    for (int k = 0; k < 3; k++)
      p[k] += c2[k];
    gilbert3(wk, p, a, b, c3);
    for (int k = 0; k < 3; k++) {
      p[k] += (a[k] - da[k]) - dc[k];
      n1[k] = -c2[k];
      n2[k] = -a3[k];
    }
    gilbert3(wk, p, n1, n2, b);
  } else {
    gilbert3(wk, p, b2, c2, a2);
    for (int k = 0; k < 3; k++)
      p[k] += b2[k];
    gilbert3(wk, p, c, a2, b3);
    for (int k = 0; k < 3; k++) {
      p[k] += (c[k] - dc[k]) - db[k];
      n1[k] = -b2[k];
      n2[k] = -c3[k];
    }
    gilbert3(wk, p, a, n1, n2);
    for (int k = 0; k < 3; k++) {
      p[k] += a[k] - (da[k] - db[k]);
      n1[k] = -c[k];
      n2[k] = -a3[k];
    }
    gilbert3(wk, p, n1, n2, b3);
    for (int k = 0; k < 3; k++) {
      p[k] += -c[k] - (db[k] - dc[k]);
      n1[k] = -b2[k];
      n3[k] = -a3[k];
    }
    gilbert3(wk, p, n1, c2, n3);
  }
}

/* sfc::get_map{1,2,3}d, nix/sfc.cpp:31-141 */
int orc_sfc_build(int32_t Cz, int32_t Cy, int32_t Cx, int32_t* chunkid, int32_t* coord)
{
  if (Cz < 1 || Cy < 1 || Cx < 1)
    return 1;
  size_t n = (size_t)Cz * Cy * Cx;
  walk_t wk = {chunkid, Cy, Cx, 0};
  int    origin[3] = {0, 0, 0};
  int    ex[3] = {Cx, 0, 0}, ey[3] = {0, Cy, 0}, ez[3] = {0, 0, Cz};
  int    nlong = (Cx != 1) + (Cy != 1) + (Cz != 1);

  if (nlong == 3) {
    if (Cx >= Cy && Cx >= Cz)
      gilbert3(&wk, origin, ex, ey, ez);
    else if (Cy >= Cx && Cy >= Cz)
      gilbert3(&wk, origin, ey, ex, ez);
    else
      gilbert3(&wk, origin, ez, ex, ey);
  } else if (nlong == 2) {
    int *eu, *ev;
    if (Cz == 1) {
      eu = ex;
      ev = ey;
    } else if (Cy == 1) {
      eu = ex;
      ev = ez;
    } else {
      eu = ey;
      ev = ez;
    }
    if (v3len(eu) >= v3len(ev))
      gilbert2(&wk, origin, eu, ev);
    else
      gilbert2(&wk, origin, ev, eu);
  } else {
    for (size_t i = 0; i < n; i++)
      chunkid[i] = (int32_t)i;
  }
  for (int iz = 0; iz < Cz; iz++)
    for (int iy = 0; iy < Cy; iy++)
      for (int ix = 0; ix < Cx; ix++) {
        int id            = chunkid[(size_t)(iz * Cy + iy) * Cx + ix];
        coord[3 * id + 0] = ix;
        coord[3 * id + 1] = iy;
        coord[3 * id + 2] = iz;
      }
  return 0;
}

static int upper_bound_d(const double* a, int n, double v)
{
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (a[mid] <= v)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

/* Balancer::assign_binarysearch, nix/balancer.cpp:71-99 */
static int assign_binarysearch(const double* load, int nc, int nr, int32_t* boundary)
{
  double* cum = (double*)malloc(sizeof(double) * (nc + 1));
  cum[0]      = 0;
  for (int i = 0; i < nc; i++)
    cum[i + 1] = cum[i] + load[i];
  double mean  = cum[nc] / nr;
  boundary[0]  = 0;
  boundary[nr] = nc;
  for (int i = 1; i < nr; i++)
    boundary[i] = upper_bound_d(cum, nc + 1, mean * i) - 1;
  free(cum);
  int ok = boundary[0] == 0 && boundary[nr] == nc;
  for (int i = 1; i < nr; i++)
    ok = ok && (boundary[i + 1] > boundary[i]);
  return ok;
}

/* Balancer::assign_smilei, nix/balancer.cpp:8-69 */
static int assign_smilei(const double* load, int nc, int nr, int32_t* boundary)
{
  double*  cum = (double*)malloc(sizeof(double) * (nc + 1));
  int32_t* old = (int32_t*)malloc(sizeof(int32_t) * (nr + 1));
  cum[0]       = 0;
  for (int i = 0; i < nc; i++)
    cum[i + 1] = cum[i] + load[i];
  memcpy(old, boundary, sizeof(int32_t) * (nr + 1));
  double mean = cum[nc] / nr;
  for (int i = 1; i < nr; i++) {
    double target = mean * i, current = cum[boundary[i]];
    if (current > target) {
      int index = boundary[i] - 1;
      while (fabs(current - target) > fabs(current - target - load[index])) {
        current -= load[index];
        index--;
      }
      boundary[i] = (index >= old[i - 1]) ? index + 1 : old[i - 1] + 1;
    } else {
      int index = boundary[i];
      while (fabs(current - target) > fabs(current - target + load[index])) {
        current += load[index];
        index++;
      }
      boundary[i] = (index < old[i + 1]) ? index : old[i + 1] - 1;
    }
  }
  int changed = memcmp(old, boundary, sizeof(int32_t) * (nr + 1)) != 0;
  free(cum);
  free(old);
  return changed;
}

/* Balancer::assign_initial, nix/balancer.cpp:101-124 */
int orc_assign_initial(const double* load, int32_t nchunk, int32_t nrank, int32_t* boundary)
{
  if (nchunk < 1 || nrank < 1 || nrank > nchunk)
    return 1;
  if (!assign_binarysearch(load, nchunk, nrank, boundary)) {
    double* uniform = (double*)malloc(sizeof(double) * nchunk);
    for (int i = 0; i < nchunk; i++)
      uniform[i] = 1.0;
    assign_binarysearch(uniform, nchunk, nrank, boundary);
    free(uniform);
    for (int iter = 0; iter < 100; iter++)
      if (!assign_smilei(load, nchunk, nrank, boundary))
        break;
  }
  return 0;
}

int orc_assign_rebalance(const double* load, int32_t nchunk, int32_t nrank, int32_t* boundary)
{
  if (nchunk < 1 || nrank < 1 || nrank > nchunk)
    return 1;
  assign_smilei(load, nchunk, nrank, boundary);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* construction                                                                               */
/* ------------------------------------------------------------------------------------------ */

static int rank_of(const orc_sim_t* s, int id)
{
  /* ChunkMap::get_rank, nix/chunkmap.cpp:109-116 */
  if (id < 0 || id >= s->nchunk_global)
    return -1;
  int r = 0;
  while (r + 1 <= s->cfg.nrank && s->boundary[r + 1] <= id)
    r++;
  return r;
}

static int neighbor_coord(const orc_sim_t* s, int coord, int delta, int dir)
{
  /* ChunkMap::get_neighbor_coord, nix/chunkmap.cpp:95-107 */
  int cdir = coord + delta;
  if (s->cfg.periodic[dir] == 1) {
    cdir = cdir >= 0 ? cdir : s->cfg.cdims[dir] - 1;
    cdir = cdir < s->cfg.cdims[dir] ? cdir : 0;
  } else {
    cdir = (cdir >= 0 && cdir < s->cfg.cdims[dir]) ? cdir : -1;
  }
  return cdir;
}

static int dir_active(const orc_sim_t* s, int dz, int dy, int dx)
{
  /* ignorable dimensions only take part with index 1: Chunk::set_index_bounds, nix/chunk.cpp:141-169 */
  if (dz == 1 && dy == 1 && dx == 1)
    return 0;
  if (!s->has_dim[0] && dz != 1)
    return 0;
  if (!s->has_dim[1] && dy != 1)
    return 0;
  if (!s->has_dim[2] && dx != 1)
    return 0;
  return 1;
}

static int region_len(const orc_sim_t* s, int axis, int dcode)
{
  return dcode == 1 ? (s->Ub[axis] - s->Lb[axis] + 1) : s->nb;
}
static int margin_lo(const orc_sim_t* s, int axis, int dcode)
{
  /* interior margin ("send_bound" of the field halo), nix/chunk.cpp:171-207 */
  return dcode == 2 ? s->Ub[axis] - s->nb + 1 : s->Lb[axis];
}
static int ghost_lo(const orc_sim_t* s, int axis, int dcode)
{
  return dcode == 0 ? s->Lb[axis] - s->nb : (dcode == 1 ? s->Lb[axis] : s->Ub[axis] + 1);
}
static int64_t region_elems(const orc_sim_t* s, int dir, int ncomp)
{
  return (int64_t)region_len(s, 0, dir / 9) * region_len(s, 1, (dir / 3) % 3) * region_len(s, 2, dir % 3) * ncomp;
}

static int cmp_msg(const void* a, const void* b)
{
  const int* x = (const int*)a;
  const int* y = (const int*)b;
  if (x[0] != y[0])
    return x[0] < y[0] ? -1 : 1;
  return (x[1] > y[1]) - (x[1] < y[1]);
}

static void build_peers(orc_sim_t* s)
{
  int nr          = s->cfg.nrank;
  s->peer_of_rank = (int*)malloc(sizeof(int) * nr);
  for (int r = 0; r < nr; r++)
    s->peer_of_rank[r] = -1;
  s->npeer = 0;
  for (int ic = 0; ic < s->nchunk; ic++)
    for (int d = 0; d < 27; d++) {
      int r = s->chunks[ic].nbrank[d];
      if (dir_active(s, d / 9, (d / 3) % 3, d % 3) && s->chunks[ic].nbid[d] >= 0 && r != s->cfg.rank &&
          s->peer_of_rank[r] < 0)
        s->peer_of_rank[r] = 0;
    }
  for (int r = 0; r < nr; r++)
    if (s->peer_of_rank[r] == 0)
      s->peer_of_rank[r] = s->npeer++;
  s->peers = (peer_t*)calloc(s->npeer > 0 ? s->npeer : 1, sizeof(peer_t));
  for (int r = 0; r < nr; r++) {
    int pi = s->peer_of_rank[r];
    if (pi < 0)
      continue;
    peer_t* p = &s->peers[pi];
    p->rank   = r;
    /* messages: key (sender chunk id, sender dir), payload (local chunk, local dir) */
    int  cap  = s->nchunk * 26;
    int* smsg = (int*)malloc(sizeof(int) * 4 * cap);
    int* rmsg = (int*)malloc(sizeof(int) * 4 * cap);
    int  ns = 0, nrv = 0;
    for (int ic = 0; ic < s->nchunk; ic++)
      for (int d = 0; d < 27; d++) {
        chunk_t* c = &s->chunks[ic];
        if (!dir_active(s, d / 9, (d / 3) % 3, d % 3) || c->nbid[d] < 0 || c->nbrank[d] != r)
          continue;
        int opp          = 26 - d;
        smsg[4 * ns + 0] = c->id;
        smsg[4 * ns + 1] = d;
        smsg[4 * ns + 2] = ic;
        smsg[4 * ns + 3] = d;
        ns++;
        rmsg[4 * nrv + 0] = c->nbid[d];
        rmsg[4 * nrv + 1] = opp;
        rmsg[4 * nrv + 2] = ic;
        rmsg[4 * nrv + 3] = d;
        nrv++;
      }
    qsort(smsg, ns, 4 * sizeof(int), cmp_msg);
    qsort(rmsg, nrv, 4 * sizeof(int), cmp_msg);
    p->nsend      = ns;
    p->nrecv      = nrv;
    p->send_chunk = (int*)malloc(sizeof(int) * (ns + 1));
    p->send_dir   = (int*)malloc(sizeof(int) * (ns + 1));
    p->recv_chunk = (int*)malloc(sizeof(int) * (nrv + 1));
    p->recv_dir   = (int*)malloc(sizeof(int) * (nrv + 1));
    for (int mode = 0; mode < 3; mode++) {
      int ncomp         = mode == 0 ? 6 : (mode == 1 ? 4 : s->cfg.Ns * 14);
      p->send_off[mode] = (int64_t*)malloc(sizeof(int64_t) * (ns + 1));
      p->recv_off[mode] = (int64_t*)malloc(sizeof(int64_t) * (nrv + 1));
      int64_t off       = 0;
      for (int m = 0; m < ns; m++) {
        p->send_off[mode][m] = off;
        off += region_elems(s, smsg[4 * m + 3], ncomp);
      }
      p->send_elems[mode] = off;
      off                 = 0;
      for (int m = 0; m < nrv; m++) {
        p->recv_off[mode][m] = off;
        off += region_elems(s, rmsg[4 * m + 3], ncomp);
      }
      p->recv_elems[mode] = off;
      p->send[mode]       = (double*)calloc(p->send_elems[mode] + 1, sizeof(double));
      p->recv[mode]       = (double*)calloc(p->recv_elems[mode] + 1, sizeof(double));
    }
    for (int m = 0; m < ns; m++) {
      p->send_chunk[m] = smsg[4 * m + 2];
      p->send_dir[m]   = smsg[4 * m + 3];
    }
    for (int m = 0; m < nrv; m++) {
      p->recv_chunk[m] = rmsg[4 * m + 2];
      p->recv_dir[m]   = rmsg[4 * m + 3];
    }
    free(smsg);
    free(rmsg);
  }
}

orc_sim_t* orc_create(const orc_config_t* cfg, const int32_t* boundary)
{
  orc_sim_t* s = (orc_sim_t*)calloc(1, sizeof(orc_sim_t));
  s->cfg       = *cfg;
  if (cfg->order < 1 || cfg->order > 4 || cfg->Ns < 1 || cfg->nrank < 1 || cfg->rank < 0 || cfg->rank >= cfg->nrank) {
    free(s);
    return NULL;
  }
  s->order  = cfg->order;
  s->is_odd = cfg->order % 2;
  s->nb     = (cfg->order + 3) / 2; /* pic/pic_chunk.cpp:195 */
  s->W      = cfg->simd_width > 0 ? cfg->simd_width : 8;
  s->del[0] = cfg->delz;
  s->del[1] = cfg->dely;
  s->del[2] = cfg->delx;
  for (int i = 0; i < 3; i++) {
    if (cfg->ndims[i] < 1 || cfg->cdims[i] < 1 || cfg->ndims[i] % cfg->cdims[i] != 0) {
      free(s);
      return NULL;
    }
    s->has_dim[i] = (cfg->ndims[i] == 1 && cfg->cdims[i] == 1) ? 0 : 1; /* nix/application.cpp:262-266 */
    s->dims[i]    = cfg->ndims[i] / cfg->cdims[i];
    s->Lb[i]      = s->nb; /* nix/chunk.cpp:134-169 */
    s->Ub[i]      = s->has_dim[i] ? s->nb + s->dims[i] - 1 : s->nb;
    s->M[i]       = s->dims[i] + 2 * s->nb;
    s->glim[i][0] = 0.0;
    s->glim[i][1] = cfg->ndims[i] * s->del[i];
  }
  s->dimension = s->has_dim[0] ? 3 : (s->has_dim[1] ? 2 : 1);
  s->Ng        = s->M[0] * s->M[1] * s->M[2];
  s->fsy       = s->Ub[2] - s->Lb[2] + 2; /* nix/xtensor_particle.hpp:231-238 */
  s->fsz       = s->fsy * (s->Ub[1] - s->Lb[1] + 2);

  int Cz = cfg->cdims[0], Cy = cfg->cdims[1], Cx = cfg->cdims[2];
  s->nchunk_global = Cz * Cy * Cx;
  s->chunkid       = (int32_t*)malloc(sizeof(int32_t) * s->nchunk_global);
  s->coord         = (int32_t*)malloc(sizeof(int32_t) * 3 * s->nchunk_global);
  orc_sfc_build(Cz, Cy, Cx, s->chunkid, s->coord);
  s->boundary = (int32_t*)malloc(sizeof(int32_t) * (cfg->nrank + 1));
  if (boundary != NULL) {
    memcpy(s->boundary, boundary, sizeof(int32_t) * (cfg->nrank + 1));
  } else {
    double* load = (double*)malloc(sizeof(double) * s->nchunk_global);
    for (int i = 0; i < s->nchunk_global; i++)
      load[i] = 1.0;
    orc_assign_initial(load, s->nchunk_global, cfg->nrank, s->boundary);
    free(load);
  }
  s->chunk_begin = s->boundary[cfg->rank];
  s->nchunk      = s->boundary[cfg->rank + 1] - s->boundary[cfg->rank];
  s->nthread     = cfg->nthread > 0 ? cfg->nthread : omp_get_max_threads();

  s->chunks = (chunk_t*)calloc(s->nchunk, sizeof(chunk_t));
  for (int ic = 0; ic < s->nchunk; ic++) {
    chunk_t* c  = &s->chunks[ic];
    c->id       = s->chunk_begin + ic;
    c->coord[0] = s->coord[3 * c->id + 2];
    c->coord[1] = s->coord[3 * c->id + 1];
    c->coord[2] = s->coord[3 * c->id + 0];
    /* ChunkVector::set_neighbors, nix/chunkvector.hpp:56-81 */
    for (int dz = -1; dz <= 1; dz++)
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          int k  = 9 * (dz + 1) + 3 * (dy + 1) + (dx + 1);
          int nz = neighbor_coord(s, c->coord[0], dz, 0);
          int ny = neighbor_coord(s, c->coord[1], dy, 1);
          int nx = neighbor_coord(s, c->coord[2], dx, 2);
          int nb = (nz >= 0 && ny >= 0 && nx >= 0) ? s->chunkid[(size_t)(nz * Cy + ny) * Cx + nx] : -1;
          c->nbid[k]   = nb;
          c->nbrank[k] = rank_of(s, nb);
        }
    /* Chunk::set_coordinate with the offsets of nix/application.cpp:291-300 */
    for (int i = 0; i < 3; i++) {
      int offset    = c->coord[i] * cfg->ndims[i] / cfg->cdims[i];
      c->lim[i][0]  = offset * s->del[i];
      c->lim[i][1]  = offset * s->del[i] + s->dims[i] * s->del[i];
    }
    c->uf = (double*)calloc((size_t)s->Ng * 6, sizeof(double));
    c->uj = (double*)calloc((size_t)s->Ng * 4, sizeof(double));
    c->ff = (double*)calloc((size_t)s->Ng * 18, sizeof(double));
    c->um = (double*)calloc((size_t)s->Ng * cfg->Ns * 14, sizeof(double));
    c->sp = (species_t*)calloc(cfg->Ns, sizeof(species_t));
    for (int is = 0; is < cfg->Ns; is++) {
      species_t* sp = &c->sp[is];
      sp->m         = 1.0;
      sp->pindex    = (int*)calloc(s->Ng + 1, sizeof(int));
      sp->pcount    = (int*)calloc((size_t)(s->Ng + 1) * s->W, sizeof(int));
    }
    for (int d = 0; d < 27; d++)
      c->out_cnt[d] = (int*)calloc(cfg->Ns, sizeof(int));
  }
  build_peers(s);
  return s;
}

void orc_destroy(orc_sim_t* s)
{
  if (s == NULL)
    return;
  for (int ic = 0; ic < s->nchunk; ic++) {
    chunk_t* c = &s->chunks[ic];
    free(c->uf);
    free(c->uj);
    free(c->ff);
    free(c->um);
    for (int is = 0; is < s->cfg.Ns; is++) {
      free(c->sp[is].xu);
      free(c->sp[is].xv);
      free(c->sp[is].gindex);
      free(c->sp[is].pindex);
      free(c->sp[is].pcount);
    }
    free(c->sp);
    for (int d = 0; d < 27; d++) {
      free(c->out_idx[d]);
      free(c->out_cnt[d]);
    }
  }
  for (int i = 0; i < s->npeer; i++) {
    peer_t* p = &s->peers[i];
    free(p->send_chunk);
    free(p->send_dir);
    free(p->recv_chunk);
    free(p->recv_dir);
    for (int m = 0; m < 3; m++) {
      free(p->send_off[m]);
      free(p->recv_off[m]);
      free(p->send[m]);
      free(p->recv[m]);
    }
    free(p->psend);
    free(p->precv);
  }
  free(s->peers);
  free(s->peer_of_rank);
  free(s->chunks);
  free(s->chunkid);
  free(s->coord);
  free(s->boundary);
  free(s);
}

int orc_num_chunks(const orc_sim_t* s) { return s->nchunk; }
int orc_num_threads(const orc_sim_t* s) { return s->nthread; }
int orc_chunk_id_begin(const orc_sim_t* s) { return s->chunk_begin; }

void orc_get_shape(const orc_sim_t* s, int32_t* shape5)
{
  shape5[0] = s->M[0];
  shape5[1] = s->M[1];
  shape5[2] = s->M[2];
  shape5[3] = s->nb;
  shape5[4] = s->Ng;
}

void orc_get_neighbors(const orc_sim_t* s, int ic, int32_t* nbid, int32_t* nbrank)
{
  for (int k = 0; k < 27; k++) {
    nbid[k]   = s->chunks[ic].nbid[k];
    nbrank[k] = s->chunks[ic].nbrank[k];
  }
}

void orc_set_species(orc_sim_t* s, int is, double q, double m)
{
  for (int ic = 0; ic < s->nchunk; ic++) {
    s->chunks[ic].sp[is].q = q;
    s->chunks[ic].sp[is].m = m;
  }
}

static double* field_ptr(const orc_sim_t* s, int ic, int which, size_t* n)
{
  const chunk_t* c = &s->chunks[ic];
  switch (which) {
  case FIELD_UF:
    *n = (size_t)s->Ng * 6;
    return c->uf;
  case FIELD_UJ:
    *n = (size_t)s->Ng * 4;
    return c->uj;
  case FIELD_FF:
    *n = (size_t)s->Ng * 18;
    return c->ff;
  default:
    *n = (size_t)s->Ng * s->cfg.Ns * 14;
    return c->um;
  }
}

void orc_set_field(orc_sim_t* s, int ic, int which, const double* in)
{
  size_t  n;
  double* p = field_ptr(s, ic, which, &n);
  memcpy(p, in, n * sizeof(double));
}

void orc_get_field(const orc_sim_t* s, int ic, int which, double* out)
{
  size_t  n;
  double* p = field_ptr(s, ic, which, &n);
  memcpy(out, p, n * sizeof(double));
}

static void species_reserve(species_t* sp, int cap)
{
  if (cap <= sp->cap)
    return;
  /* Particle::round_up_alloc, nix/particle.hpp:146-153 */
  cap        = ((cap + ALLOC_UNIT) / ALLOC_UNIT) * ALLOC_UNIT;
  sp->xu     = (double*)realloc(sp->xu, sizeof(double) * NC * (size_t)cap);
  sp->xv     = (double*)realloc(sp->xv, sizeof(double) * NC * (size_t)cap);
  sp->gindex = (int*)realloc(sp->gindex, sizeof(int) * (size_t)cap);
  memset(sp->xu + (size_t)NC * sp->cap, 0, sizeof(double) * NC * (size_t)(cap - sp->cap));
  memset(sp->xv + (size_t)NC * sp->cap, 0, sizeof(double) * NC * (size_t)(cap - sp->cap));
  sp->cap = cap;
}

void orc_set_particles(orc_sim_t* s, int ic, int is, const double* xu, int np, int np_alloc)
{
  species_t* sp = &s->chunks[ic].sp[is];
  species_reserve(sp, np_alloc > np ? np_alloc : np);
  memcpy(sp->xu, xu, sizeof(double) * NC * (size_t)np);
  sp->np = np;
}

int orc_get_np(const orc_sim_t* s, int ic, int is) { return s->chunks[ic].sp[is].np; }

void orc_get_particles(const orc_sim_t* s, int ic, int is, int which, int n, double* out)
{
  const species_t* sp = &s->chunks[ic].sp[is];
  memcpy(out, which == 0 ? sp->xu : sp->xv, sizeof(double) * NC * (size_t)n);
}

void orc_get_pindex(const orc_sim_t* s, int ic, int is, int32_t* out)
{
  memcpy(out, s->chunks[ic].sp[is].pindex, sizeof(int) * (s->Ng + 1));
}

void orc_get_gindex(const orc_sim_t* s, int ic, int is, int n, int32_t* out)
{
  memcpy(out, s->chunks[ic].sp[is].gindex, sizeof(int) * (size_t)n);
}

/* ------------------------------------------------------------------------------------------ */
/* Maxwell solver: pic/engine/maxwell.hpp                                                     */
/* ------------------------------------------------------------------------------------------ */

#define CELL(s, iz, iy, ix) (((size_t)(iz) * (s)->M[1] + (iy)) * (s)->M[2] + (ix))
#define UF(c, s, iz, iy, ix, k) ((c)->uf[CELL(s, iz, iy, ix) * 6 + (k)])
#define UJ(c, s, iz, iy, ix, k) ((c)->uj[CELL(s, iz, iy, ix) * 4 + (k)])
#define FF(c, s, iz, iy, ix, t, k) ((c)->ff[CELL(s, iz, iy, ix) * 18 + (t) * 6 + (k)])

/* loop range of an axis: whole padded extent, or the single interior plane of an ignorable axis
 * (`int iz = lbz;` in the 1-D/2-D routines, maxwell.hpp:85-146, 231-290) */
static void axis_range(const orc_sim_t* s, int a, int* lo, int* hi)
{
  if (s->has_dim[a]) {
    *lo = 0;
    *hi = s->M[a] - 1;
  } else {
    *lo = *hi = s->Lb[a];
  }
}

/* init_friedman visits the whole array in every dimensionality, maxwell.hpp:44-61 */
static void chunk_init_friedman(const orc_sim_t* s, chunk_t* c)
{
  for (size_t cell = 0; cell < (size_t)s->Ng; cell++)
    for (int k = 0; k < 3; k++) {
      double e               = c->uf[cell * 6 + k];
      c->ff[cell * 18 + k]      = e;
      c->ff[cell * 18 + 6 + k]  = e;
      c->ff[cell * 18 + 12 + k] = e;
    }
}

/* push_bfd_{1,2,3}d, maxwell.hpp:148-201, 292-348, 442-496 */
static void chunk_push_bfd(const orc_sim_t* s, chunk_t* c, double delt)
{
  const double theta = s->cfg.friedman;
  const double A     = 1 + 0.5 * theta;
  const double B     = -theta * (1 - 0.5 * theta);
  const double C     = 0.5 * theta * (1 - theta) * (1 - theta);
  const double cflx = s->cfg.cc * delt / s->del[2], cfly = s->cfg.cc * delt / s->del[1],
               cflz = s->cfg.cc * delt / s->del[0];
  const int hz = s->has_dim[0], hy = s->has_dim[1];
  int       z0, z1, y0, y1, x0, x1;
  axis_range(s, 0, &z0, &z1);
  axis_range(s, 1, &y0, &y1);
  axis_range(s, 2, &x0, &x1);

  for (int iz = z0; iz <= z1; iz++)
    for (int iy = y0; iy <= y1; iy++)
      for (int ix = x0; ix <= x1; ix++)
        for (int k = 0; k < 3; k++)
          FF(c, s, iz, iy, ix, 0, k) =
              A * UF(c, s, iz, iy, ix, k) + B * FF(c, s, iz, iy, ix, 1, k) + C * FF(c, s, iz, iy, ix, 2, k);

  /* Bx: needs y or z derivatives */
  if (hy) {
    for (int iz = z0 + hz; iz <= z1; iz++)
      for (int iy = y0 + 1; iy <= y1; iy++)
        for (int ix = x0; ix <= x1; ix++) {
          double v = (-cfly) * (FF(c, s, iz, iy, ix, 0, 2) - FF(c, s, iz, iy - 1, ix, 0, 2));
          if (hz)
            v = v + (+cflz) * (FF(c, s, iz, iy, ix, 0, 1) - FF(c, s, iz - 1, iy, ix, 0, 1));
          UF(c, s, iz, iy, ix, 3) += v;
        }
  }
  /* By */
  for (int iz = z0 + hz; iz <= z1; iz++)
    for (int iy = y0; iy <= y1; iy++)
      for (int ix = x0 + 1; ix <= x1; ix++) {
        double v = (+cflx) * (FF(c, s, iz, iy, ix, 0, 2) - FF(c, s, iz, iy, ix - 1, 0, 2));
        if (hz)
          v = (-cflz) * (FF(c, s, iz, iy, ix, 0, 0) - FF(c, s, iz - 1, iy, ix, 0, 0)) + v;
        UF(c, s, iz, iy, ix, 4) += v;
      }
  /* Bz */
  for (int iz = z0; iz <= z1; iz++)
    for (int iy = y0 + hy; iy <= y1; iy++)
      for (int ix = x0 + 1; ix <= x1; ix++) {
        double v = (-cflx) * (FF(c, s, iz, iy, ix, 0, 1) - FF(c, s, iz, iy, ix - 1, 0, 1));
        if (hy)
          v = v + (+cfly) * (FF(c, s, iz, iy, ix, 0, 0) - FF(c, s, iz, iy - 1, ix, 0, 0));
        UF(c, s, iz, iy, ix, 5) += v;
      }
}

/* push_efd_{1,2,3}d, maxwell.hpp:85-146, 231-290, 383-440 */
static void chunk_push_efd(const orc_sim_t* s, chunk_t* c, double delt)
{
  const double theta = s->cfg.friedman;
  const double cflx = s->cfg.cc * delt / s->del[2], cfly = s->cfg.cc * delt / s->del[1],
               cflz = s->cfg.cc * delt / s->del[0];
  const int hz = s->has_dim[0], hy = s->has_dim[1];
  int       z0, z1, y0, y1, x0, x1;
  axis_range(s, 0, &z0, &z1);
  axis_range(s, 1, &y0, &y1);
  axis_range(s, 2, &x0, &x1);

  for (int iz = z0; iz <= z1; iz++)
    for (int iy = y0; iy <= y1; iy++)
      for (int ix = x0; ix <= x1; ix++)
        for (int k = 0; k < 3; k++) {
          FF(c, s, iz, iy, ix, 2, k) = FF(c, s, iz, iy, ix, 1, k) + theta * FF(c, s, iz, iy, ix, 2, k);
          FF(c, s, iz, iy, ix, 1, k) = UF(c, s, iz, iy, ix, k);
        }

  /* Ex */
  for (int iz = z0; iz <= z1 - hz; iz++)
    for (int iy = y0; iy <= y1 - hy; iy++)
      for (int ix = x0; ix <= x1; ix++) {
        if (hy && hz) {
          UF(c, s, iz, iy, ix, 0) += (+cfly) * (UF(c, s, iz, iy + 1, ix, 5) - UF(c, s, iz, iy, ix, 5)) +
                                     (-cflz) * (UF(c, s, iz + 1, iy, ix, 4) - UF(c, s, iz, iy, ix, 4)) -
                                     delt * UJ(c, s, iz, iy, ix, 1);
        } else if (hy) {
          UF(c, s, iz, iy, ix, 0) +=
              (+cfly) * (UF(c, s, iz, iy + 1, ix, 5) - UF(c, s, iz, iy, ix, 5)) - delt * UJ(c, s, iz, iy, ix, 1);
        } else {
          UF(c, s, iz, iy, ix, 0) += -delt * UJ(c, s, iz, iy, ix, 1);
        }
      }
  /* Ey */
  for (int iz = z0; iz <= z1 - hz; iz++)
    for (int iy = y0; iy <= y1; iy++)
      for (int ix = x0; ix <= x1 - 1; ix++) {
        if (hz) {
          UF(c, s, iz, iy, ix, 1) += (+cflz) * (UF(c, s, iz + 1, iy, ix, 3) - UF(c, s, iz, iy, ix, 3)) +
                                     (-cflx) * (UF(c, s, iz, iy, ix + 1, 5) - UF(c, s, iz, iy, ix, 5)) -
                                     delt * UJ(c, s, iz, iy, ix, 2);
        } else {
          UF(c, s, iz, iy, ix, 1) +=
              (-cflx) * (UF(c, s, iz, iy, ix + 1, 5) - UF(c, s, iz, iy, ix, 5)) - delt * UJ(c, s, iz, iy, ix, 2);
        }
      }
  /* Ez */
  for (int iz = z0; iz <= z1; iz++)
    for (int iy = y0; iy <= y1 - hy; iy++)
      for (int ix = x0; ix <= x1 - 1; ix++) {
        if (hy) {
          UF(c, s, iz, iy, ix, 2) += (+cflx) * (UF(c, s, iz, iy, ix + 1, 4) - UF(c, s, iz, iy, ix, 4)) +
                                     (-cfly) * (UF(c, s, iz, iy + 1, ix, 3) - UF(c, s, iz, iy, ix, 3)) -
                                     delt * UJ(c, s, iz, iy, ix, 3);
        } else {
          UF(c, s, iz, iy, ix, 2) +=
              (+cflx) * (UF(c, s, iz, iy, ix + 1, 4) - UF(c, s, iz, iy, ix, 4)) - delt * UJ(c, s, iz, iy, ix, 3);
        }
      }
}

/* Maxwell::get_diverror + get_diverror_{1,2,3}d, pic/pic_engine.hpp:36-89, maxwell.hpp:63-83,203-229,350-381 */
void orc_get_diverror(const orc_sim_t* s, int ic, double* efd, double* bfd)
{
  const chunk_t* c = &s->chunks[ic];
  int            lo[3], hi[3];
  for (int a = 0; a < 3; a++) {
    lo[a] = s->Lb[a];
    hi[a] = s->Ub[a];
  }
  /* skip the margin next to a physical boundary (MPI_PROC_NULL neighbour) */
  if (c->nbid[9 + 3 + 0] < 0)
    lo[2] += s->nb;
  if (c->nbid[9 + 3 + 2] < 0)
    hi[2] -= s->nb;
  if (c->nbid[9 + 0 + 1] < 0)
    lo[1] += s->nb;
  if (c->nbid[9 + 6 + 1] < 0)
    hi[1] -= s->nb;
  if (c->nbid[0 + 3 + 1] < 0)
    lo[0] += s->nb;
  if (c->nbid[18 + 3 + 1] < 0)
    hi[0] -= s->nb;
  for (int a = 0; a < 2; a++)
    if (!s->has_dim[a])
      lo[a] = hi[a] = s->Lb[a];
  const double rdx = 1 / s->del[2], rdy = 1 / s->del[1], rdz = 1 / s->del[0];
  double       e = 0, b = 0;
  for (int iz = lo[0]; iz <= hi[0]; iz++)
    for (int iy = lo[1]; iy <= hi[1]; iy++)
      for (int ix = lo[2]; ix <= hi[2]; ix++) {
        double dive = (UF(c, s, iz, iy, ix + 1, 0) - UF(c, s, iz, iy, ix, 0)) * rdx;
        double divb = (UF(c, s, iz, iy, ix, 3) - UF(c, s, iz, iy, ix - 1, 3)) * rdx;
        if (s->has_dim[1]) {
          dive = dive + (UF(c, s, iz, iy + 1, ix, 1) - UF(c, s, iz, iy, ix, 1)) * rdy;
          divb = divb + (UF(c, s, iz, iy, ix, 4) - UF(c, s, iz, iy - 1, ix, 4)) * rdy;
        }
        if (s->has_dim[0]) {
          dive = dive + (UF(c, s, iz + 1, iy, ix, 2) - UF(c, s, iz, iy, ix, 2)) * rdz;
          divb = divb + (UF(c, s, iz, iy, ix, 5) - UF(c, s, iz - 1, iy, ix, 5)) * rdz;
        }
        e += dive - UJ(c, s, iz, iy, ix, 0);
        b += divb;
      }
  *efd = e;
  *bfd = b;
}

/* PicChunk::get_energy, pic/pic_chunk.cpp:407-439 */
void orc_get_energy(const orc_sim_t* s, int ic, double* efd, double* bfd, double* particle)
{
  const chunk_t* c  = &s->chunks[ic];
  const int      Ns = s->cfg.Ns;
  const double   cc = s->cfg.cc;
  double         e = 0, b = 0;
  for (int is = 0; is < Ns; is++)
    particle[is] = 0;
  for (int iz = s->Lb[0]; iz <= s->Ub[0]; iz++)
    for (int iy = s->Lb[1]; iy <= s->Ub[1]; iy++)
      for (int ix = s->Lb[2]; ix <= s->Ub[2]; ix++) {
        const double* f = &c->uf[CELL(s, iz, iy, ix) * 6];
        e += 0.5 * (f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
        b += 0.5 * (f[3] * f[3] + f[4] * f[4] + f[5] * f[5]);
      }
  for (int iz = s->Lb[0]; iz <= s->Ub[0]; iz++)
    for (int iy = s->Lb[1]; iy <= s->Ub[1]; iy++)
      for (int ix = s->Lb[2]; ix <= s->Ub[2]; ix++)
        for (int is = 0; is < Ns; is++) {
          const double* m = &c->um[(CELL(s, iz, iy, ix) * Ns + is) * 14];
          particle[is] += m[4] * cc - m[0] * cc * cc;
        }
  *efd = e;
  *bfd = b;
}

/* ------------------------------------------------------------------------------------------ */
/* particle primitives: nix/primitives.hpp                                                    */
/* ------------------------------------------------------------------------------------------ */

static int digitize(double x, double xmin, double rdx) { return (int)floor((x - xmin) * rdx); } /* :45-58 */

/* shape_mc<Order>, nix/primitives.hpp:255-329 */
static void shape_mc(int order, double x, double X, double rdx, double* s)
{
  const double delta = (x - X) * rdx;
  if (order == 1) {
    s[0] = 1 - delta;
    s[1] = delta;
  } else if (order == 2) {
    const double w1 = 0.5 - delta, w2 = 0.5 + delta;
    s[0] = 0.50 * w1 * w1;
    s[1] = 0.75 - delta * delta;
    s[2] = 0.50 * w2 * w2;
  } else if (order == 3) {
    const double a  = 1 / 6.0;
    const double w1 = delta, w2 = 1 - delta;
    const double w1_2 = w1 * w1, w2_2 = w2 * w2;
    const double w1_3 = w1_2 * w1, w2_3 = w2_2 * w2;
    s[0] = a * w2_3;
    s[1] = a * (4 - 6 * w1_2 + 3 * w1_3);
    s[2] = a * (4 - 6 * w2_2 + 3 * w2_3);
    s[3] = a * w1_3;
  } else {
    const double a = 1 / 384.0, b = 1 / 96.0, c = 115 / 192.0, d = 1 / 8.0;
    const double w1 = 1 + delta, w2 = 1 - delta, w3 = 1 + delta * 2, w4 = 1 - delta * 2;
    const double w0_2 = delta * delta;
    const double w1_2 = w1 * w1, w2_2 = w2 * w2;
    const double w1_3 = w1_2 * w1, w2_3 = w2_2 * w2;
    const double w1_4 = w1_3 * w1, w2_4 = w2_3 * w2;
    const double w3_4 = w3 * w3 * w3 * w3, w4_4 = w4 * w4 * w4 * w4;
    s[0] = a * w4_4;
    s[1] = b * (55 + 20 * w1 - 120 * w1_2 + 80 * w1_3 - 16 * w1_4);
    s[2] = c + d * w0_2 * (2 * w0_2 - 5);
    s[3] = b * (55 + 20 * w2 - 120 * w2_2 + 80 * w2_3 - 16 * w2_4);
    s[4] = a * w3_4;
  }
}

/* shape_wt<Order>, nix/primitives.hpp:331-495 (Lu et al. 2020); dt = c*delt/dx */
static void shape_wt(int order, double x, double X, double rdx, double dt, double rdt, double* s)
{
  const double delta = (x - X) * rdx;
  if (order == 1) {
    double ss = 0.25 * rdt * (1 + 2 * dt - 2 * delta);
    ss        = fmin(1.0, fmax(0.0, ss));
    s[0]      = ss;
    s[1]      = 1 - ss;
    return;
  }
  double t1, t2, t3, t4;
  double v1[5] = {0}, v2[5] = {0}, v3[5] = {0};
  int    n = order + 1;
  if (order == 2) {
    t1 = delta < -dt ? 1.0 : 0.0;
    t3 = delta < +dt ? 1.0 : 0.0;
    const double w0 = fabs(delta), w1 = dt - delta, w2 = dt + delta;
    v1[0] = w0;
    v1[1] = 1 - w0;
    v1[2] = 0;
    v2[0] = 0.25 * rdt * w1 * w1;
    v2[1] = 0.50 * rdt * (dt * (2 - dt) - w0 * w0);
    v2[2] = 0.25 * rdt * w2 * w2;
    v3[0] = v1[2];
    v3[1] = v1[1];
    v3[2] = v1[0];
  } else if (order == 3) {
    const double a = 1 / 96.0, b = 1 / 24.0, c = 1 / 12.0;
    const double adt = a * rdt;
    t1 = delta < 0.5 - dt ? 1.0 : 0.0;
    t3 = delta < 0.5 + dt ? 1.0 : 0.0;
    const double w0 = delta, w1 = 1 - delta, w3 = 1 - 2 * delta, w4 = 1 + 2 * delta;
    const double w5 = 2 * dt + w3, w6 = 2 * dt - w3, w7 = 3 - 2 * delta;
    const double w0_2 = w0 * w0, w1_2 = w1 * w1, w3_2 = w3 * w3, w3_3 = w3_2 * w3, w4_2 = w4 * w4;
    const double w5_3 = w5 * w5 * w5, w6_3 = w6 * w6 * w6, w7_2 = w7 * w7;
    const double dt_2 = dt * dt, dt_3 = dt_2 * dt, dt_2_4 = 4 * dt_2;
    const double s_2_odd  = adt * (-8 * dt_3 - 6 * dt * w3_2);
    const double s_2_even = adt * (-36 * dt_2 * w3 - 3 * w3_3);
    v1[0] = b * (dt_2_4 + 3 * w3_2);
    v1[1] = c * (9 - dt_2_4 - 12 * w0_2);
    v1[2] = b * (dt_2_4 + 3 * w4_2);
    v1[3] = 0;
    v2[0] = adt * w5_3;
    v2[1] = s_2_odd + s_2_even + w1;
    v2[2] = s_2_odd - s_2_even + w0;
    v2[3] = adt * w6_3;
    v3[0] = 0;
    v3[1] = b * (dt_2_4 + 3 * w7_2);
    v3[2] = c * (9 - dt_2_4 - 12 * w1_2);
    v3[3] = b * (dt_2_4 + 3 * w3_2);
  } else {
    const double a = 1 / 48.0, b = 1 / 24.0, c = 1 / 12.0, d = 1 / 6.0;
    const double adt = a * rdt, bdt = b * rdt, cdt = c * rdt;
    t1 = delta < -dt ? 1.0 : 0.0;
    t3 = delta < +dt ? 1.0 : 0.0;
    const double w0 = fabs(delta), w1 = 1 - w0, w2 = 1 - delta, w3 = 1 + delta, w4 = dt - delta, w5 = dt + delta;
    const double w0_2 = w0 * w0, w0_3 = w0_2 * w0, w0_4 = w0_3 * w0;
    const double w1_2 = w1 * w1, w1_3 = w1_2 * w1;
    const double w2_3 = w2 * w2 * w2, w3_3 = w3 * w3 * w3;
    const double w4_4 = w4 * w4 * w4 * w4, w5_4 = w5 * w5 * w5 * w5;
    const double dt_2 = dt * dt, dt_3 = dt_2 * dt, dt_4 = dt_3 * dt;
    const double ss1 = -dt_4 - 6 * w0_2 * dt_2 - w0_4;
    const double ss2 = 3 * dt_4 - 8 * dt_3 + 18 * w0_2 * dt_2 + (16 - 24 * w0_2) * dt + 3 * w0_4;
    v1[0] = d * w0 * (w0_2 + dt_2);
    v1[1] = d * (4 - 6 * w1_2 + 3 * w1_3 + (1 - 3 * w0) * dt_2);
    v1[2] = d * (4 - 6 * w0_2 + 3 * w0_3 - (2 - 3 * w0) * dt_2);
    v1[3] = d * w1 * (w1_2 + dt_2);
    v1[4] = 0;
    v2[0] = adt * w4_4;
    v2[1] = cdt * (ss1 + 2 * dt_3 * w3 + 2 * dt * (-6 * delta + w3_3));
    v2[2] = bdt * ss2;
    v2[3] = cdt * (ss1 + 2 * dt_3 * w2 + 2 * dt * (+6 * delta + w2_3));
    v2[4] = adt * w5_4;
    for (int k = 0; k < 5; k++)
      v3[k] = v1[4 - k];
  }
  t2 = 1 - t1;
  t4 = 1 - t3;
  for (int k = 0; k < n; k++)
    s[k] = v1[k] * t1 + v2[k] * t2 * t3 + v3[k] * t4;
}

/* push_boris, nix/primitives.hpp:164-189 */
static void push_boris(double u[3], double ex, double ey, double ez, double bx, double by, double bz, double cc)
{
  u[0] += ex;
  u[1] += ey;
  u[2] += ez;
  const double gm = 1 / sqrt(cc * cc + u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  bx *= gm;
  by *= gm;
  bz *= gm;
  const double bb = 2.0 / (1.0 + bx * bx + by * by + bz * bz);
  const double vx = u[0] + (u[1] * bz - u[2] * by);
  const double vy = u[1] + (u[2] * bx - u[0] * bz);
  const double vz = u[2] + (u[0] * by - u[1] * bx);
  u[0] += (vy * bz - vz * by) * bb + ex;
  u[1] += (vz * bx - vx * bz) * bb + ey;
  u[2] += (vx * by - vy * bx) * bb + ez;
}

/* push_vay, nix/primitives.hpp:191-221 */
static void push_vay(double u[3], double ex, double ey, double ez, double bx, double by, double bz, double cc)
{
  double       gm = 1 / sqrt(cc * cc + u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  const double vx = u[0] + 2 * ex + gm * (u[1] * bz - u[2] * by);
  const double vy = u[1] + 2 * ey + gm * (u[2] * bx - u[0] * bz);
  const double vz = u[2] + 2 * ez + gm * (u[0] * by - u[1] * bx);
  gm              = (cc * cc + vx * vx + vy * vy + vz * vz);
  double bb       = bx * bx + by * by + bz * bz;
  double bu       = bx * vx + by * vy + bz * vz;
  const double xx = gm - bb, yy = bb + bu * bu;
  gm = 1 / sqrt(0.5 * (xx + sqrt(xx * xx + 4 * yy)));
  bx *= gm;
  by *= gm;
  bz *= gm;
  bu   = bx * vx + by * vy + bz * vz;
  bb   = 1.0 / (1.0 + bx * bx + by * by + bz * bz);
  u[0] = (vx + bu * bx + (vy * bz - vz * by)) * bb;
  u[1] = (vy + bu * by + (vz * bx - vx * bz)) * bb;
  u[2] = (vz + bu * bz + (vx * by - vy * bx)) * bb;
}

/* push_higuera_cary, nix/primitives.hpp:223-253 */
static void push_hc(double u[3], double ex, double ey, double ez, double bx, double by, double bz, double cc)
{
  u[0] += ex;
  u[1] += ey;
  u[2] += ez;
  double       gm = cc * cc + u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
  double       bb = bx * bx + by * by + bz * bz;
  const double bu = bx * u[0] + by * u[1] + bz * u[2];
  const double xx = gm - bb, yy = bb + bu * bu;
  gm = 1 / sqrt(0.5 * (xx + sqrt(xx * xx + 4 * yy)));
  bx *= gm;
  by *= gm;
  bz *= gm;
  bb = 2.0 / (1.0 + bx * bx + by * by + bz * bz);
  const double vx = u[0] + (u[1] * bz - u[2] * by);
  const double vy = u[1] + (u[2] * bx - u[0] * bz);
  const double vz = u[2] + (u[0] * by - u[1] * bx);
  u[0] += (vy * bz - vz * by) * bb + ex;
  u[1] += (vz * bx - vx * bz) * bb + ey;
  u[2] += (vx * by - vy * bx) * bb + ez;
}

/* ------------------------------------------------------------------------------------------ */
/* velocity push: pic/engine/velocity.hpp (scalar path :101-136, weights :222-359, push :361-450) */
/* ------------------------------------------------------------------------------------------ */

static void axis_weights(const orc_sim_t* s, int a, double x, double xmin, double delt, double* wi, double* wh,
                         int* i0, int* h0)
{
  const int    order = s->order;
  const double dx = s->del[a], rdx = 1 / dx;
  const double ximin  = xmin + 0.5 * dx * s->is_odd;
  const double xhmin  = xmin + 0.5 * dx * s->is_odd - 0.5 * dx;
  const double xigrid = xmin + 0.5 * dx;
  const double xhgrid = xmin;
  *i0 = digitize(x, ximin, rdx);
  *h0 = digitize(x, xhmin, rdx);
  const double xig = xigrid + *i0 * dx;
  const double xhg = xhgrid + *h0 * dx;
  if (s->cfg.interp == 0) {
    shape_mc(order, x, xig, rdx, wi);
  } else {
    const double cfl = s->cfg.cc * delt / dx;
    shape_wt(order, x, xig, rdx, cfl, 1 / cfl, wi);
  }
  shape_mc(order, x, xhg, rdx, wh);
  *i0 += s->Lb[a] - order / 2;
  *h0 += s->Lb[a] - order / 2;
}

/* interp{1,2,3}d, nix/interp.hpp:14-113: x innermost, then y, then z, times dt */
static double interp(const orc_sim_t* s, const chunk_t* c, int iz0, int iy0, int ix0, int k, const double* wz,
                     const double* wy, const double* wx, double dt)
{
  const int N = s->order + 1;
  if (s->dimension == 1) {
    double rx = 0;
    for (int jx = 0; jx < N; jx++)
      rx += UF(c, s, iz0, iy0, ix0 + jx, k) * wx[jx];
    return rx * dt;
  }
  if (s->dimension == 2) {
    double ry = 0;
    for (int jy = 0; jy < N; jy++) {
      double rx = 0;
      for (int jx = 0; jx < N; jx++)
        rx += UF(c, s, iz0, iy0 + jy, ix0 + jx, k) * wx[jx];
      ry += rx * wy[jy];
    }
    return ry * dt;
  }
  double rz = 0;
  for (int jz = 0; jz < N; jz++) {
    double ry = 0;
    for (int jy = 0; jy < N; jy++) {
      double rx = 0;
      for (int jx = 0; jx < N; jx++)
        rx += UF(c, s, iz0 + jz, iy0 + jy, ix0 + jx, k) * wx[jx];
      ry += rx * wy[jy];
    }
    rz += ry * wz[jz];
  }
  return rz * dt;
}

static void chunk_push_velocity(const orc_sim_t* s, chunk_t* c, double delt)
{
  for (int is = 0; is < s->cfg.Ns; is++) {
    species_t*   sp   = &c->sp[is];
    const double qmdt = 0.5 * sp->q / sp->m * delt;
    for (int ip = 0; ip < sp->np; ip++) {
      double* xu = &sp->xu[(size_t)NC * ip];
      double  wix[5], whx[5], wiy[5], why[5], wiz[5], whz[5];
      int     ix0, hx0, iy0 = s->Lb[1], hy0 = s->Lb[1], iz0 = s->Lb[0], hz0 = s->Lb[0];
      axis_weights(s, 2, xu[0], c->lim[2][0], delt, wix, whx, &ix0, &hx0);
      if (s->dimension >= 2)
        axis_weights(s, 1, xu[1], c->lim[1][0], delt, wiy, why, &iy0, &hy0);
      if (s->dimension >= 3)
        axis_weights(s, 0, xu[2], c->lim[0][0], delt, wiz, whz, &iz0, &hz0);
      /* Yee staggering, velocity.hpp:379-384, 410-415, 442-447 */
      double ex = interp(s, c, iz0, iy0, hx0, 0, wiz, wiy, whx, qmdt);
      double ey = interp(s, c, iz0, hy0, ix0, 1, wiz, why, wix, qmdt);
      double ez = interp(s, c, hz0, iy0, ix0, 2, whz, wiy, wix, qmdt);
      double bx = interp(s, c, hz0, hy0, ix0, 3, whz, why, wix, qmdt);
      double by = interp(s, c, hz0, iy0, hx0, 4, whz, wiy, whx, qmdt);
      double bz = interp(s, c, iz0, hy0, hx0, 5, wiz, why, whx, qmdt);
      if (s->cfg.pusher == 0)
        push_boris(xu + 3, ex, ey, ez, bx, by, bz, s->cfg.cc);
      else if (s->cfg.pusher == 1)
        push_vay(xu + 3, ex, ey, ez, bx, by, bz, s->cfg.cc);
      else
        push_hc(xu + 3, ex, ey, ez, bx, by, bz, s->cfg.cc);
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* particle container: nix/xtensor_particle.hpp                                               */
/* ------------------------------------------------------------------------------------------ */

/* XtensorParticle::count, nix/xtensor_particle.hpp:324-357 */
static void species_count(const orc_sim_t* s, const chunk_t* c, species_t* sp, int lbp, int ubp, int reset)
{
  const int    W    = s->W;
  const double half = 0.5 * s->is_odd;
  const double xoff = c->lim[2][0] - half * s->del[2], yoff = c->lim[1][0] - half * s->del[1],
               zoff = c->lim[0][0] - half * s->del[0];
  const double rdx = 1 / s->del[2], rdy = 1 / s->del[1], rdz = 1 / s->del[0];
  if (reset)
    memset(sp->pcount, 0, sizeof(int) * (size_t)(s->Ng + 1) * W);
  for (int ip = lbp; ip <= ubp; ip++) {
    const double* p  = &sp->xu[(size_t)NC * ip];
    int           ix = s->has_dim[2] ? digitize(p[0], xoff, rdx) : 0;
    int           iy = s->has_dim[1] ? digitize(p[1], yoff, rdy) : 0;
    int           iz = s->has_dim[0] ? digitize(p[2], zoff, rdz) : 0;
    int           ii = iz * s->fsz + iy * s->fsy + ix;
    ii = (s->has_dim[2] && (p[0] < c->lim[2][0] || p[0] >= c->lim[2][1])) ? s->Ng : ii;
    ii = (s->has_dim[1] && (p[1] < c->lim[1][0] || p[1] >= c->lim[1][1])) ? s->Ng : ii;
    ii = (s->has_dim[0] && (p[2] < c->lim[0][0] || p[2] >= c->lim[0][1])) ? s->Ng : ii;
    sp->gindex[ip] = ii; /* increment(), :246-252 */
    sp->pcount[(size_t)ii * W + ip % W]++;
  }
}

/* XtensorParticle::sort, nix/xtensor_particle.hpp:260-321: stable in (cell, ip % W, ip) */
static void species_sort(const orc_sim_t* s, species_t* sp)
{
  const int W = s->W, Ng = s->Ng;
  int*      pc = sp->pcount;
  for (int ii = 0; ii < Ng + 1; ii++)
    for (int jj = 0; jj < W - 1; jj++)
      pc[(size_t)ii * W + jj + 1] += pc[(size_t)ii * W + jj];
  for (int ii = 0; ii < Ng; ii++)
    for (int jj = 0; jj < W; jj++)
      pc[(size_t)(ii + 1) * W + jj] += pc[(size_t)ii * W + W - 1];
  sp->pindex[0] = 0;
  for (int ii = 0; ii < Ng; ii++)
    sp->pindex[ii + 1] = pc[(size_t)ii * W + W - 1];
  for (int ii = 0; ii < Ng + 1; ii++)
    for (int jj = W - 1; jj > 0; jj--)
      pc[(size_t)ii * W + jj] = pc[(size_t)ii * W + jj - 1];
  for (int ii = 0; ii < Ng + 1; ii++)
    pc[(size_t)ii * W] = sp->pindex[ii];
  for (int ip = 0; ip < sp->np; ip++) {
    int ii = sp->gindex[ip], jj = ip % W;
    int jp = pc[(size_t)ii * W + jj];
    if (jp < sp->cap) /* the discarded tail (key Ng) may run past the live range */
      memcpy(&sp->xv[(size_t)NC * jp], &sp->xu[(size_t)NC * ip], sizeof(double) * NC);
    pc[(size_t)ii * W + jj]++;
  }
  double* t = sp->xu;
  sp->xu    = sp->xv;
  sp->xv    = t;
  sp->np    = sp->pindex[Ng];
}

/* Position: pic/engine/position.hpp:55-130 + pic_engine::Position::set_boundary (pic_engine.hpp:292-303) */
static void chunk_push_position(const orc_sim_t* s, chunk_t* c, double delt)
{
  const double rc = 1 / s->cfg.cc;
  for (int is = 0; is < s->cfg.Ns; is++) {
    species_t* sp = &c->sp[is];
    for (int ip = 0; ip < sp->np; ip++) {
      double* xu = &sp->xu[(size_t)NC * ip];
      double* xv = &sp->xv[(size_t)NC * ip];
      memcpy(xv, xu, sizeof(double) * NC);
      const double gm = sqrt(1 + (xu[3] * xu[3] + xu[4] * xu[4] + xu[5] * xu[5]) * rc * rc);
      const double dt = delt / gm;
      xu[0] += xu[3] * dt;
      xu[1] += xu[4] * dt;
      xu[2] += xu[5] * dt;
    }
    species_count(s, c, sp, 0, sp->np - 1, 1);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Esirkepov deposit: pic/engine/current.hpp (scalar path :59-98, local{1,2,3}d :207-395),     */
/* nix/esirkepov.hpp (ro/ds/jx/jy/jz :18-237, shift_weights :240-275), append_current :678-834 */
/* ------------------------------------------------------------------------------------------ */

static int esirkepov_axis(const orc_sim_t* s, int a, double x0, double x1, double xmin, double* s0, double* s1)
{
  const int    S = s->order + 3;
  const double dx = s->del[a], rdx = 1 / dx;
  const double ximin = xmin + 0.5 * dx * s->is_odd;
  const double xgrid = xmin + 0.5 * dx;
  for (int j = 0; j < S; j++)
    s0[j] = s1[j] = 0;
  const int i0 = digitize(x0, ximin, rdx);
  shape_mc(s->order, x0, xgrid + i0 * dx, rdx, s0 + 1);
  const int i1 = digitize(x1, ximin, rdx);
  shape_mc(s->order, x1, xgrid + i1 * dx, rdx, s1 + 1);
  const int shift = i1 - i0;
  if (shift < 0) {
    for (int j = 0; j < S - 1; j++)
      s1[j] = s1[j + 1];
    s1[S - 1] = 0;
  } else if (shift > 0) {
    for (int j = S - 1; j > 0; j--)
      s1[j] = s1[j - 1];
    s1[0] = 0;
  }
  return i0 + s->Lb[a] - s->order / 2 - 1;
}

static void chunk_deposit_current(const orc_sim_t* s, chunk_t* c, double delt)
{
  const int    S = s->order + 3, dim = s->dimension;
  const double A = 1.0 / 2, B = 1.0 / 3;
  memset(c->uj, 0, sizeof(double) * (size_t)s->Ng * 4);

  for (int is = 0; is < s->cfg.Ns; is++) {
    species_t*   sp = &c->sp[is];
    const double q  = sp->q;
    for (int ip = 0; ip < sp->np; ip++) {
      const double* xv = &sp->xv[(size_t)NC * ip];
      const double* xu = &sp->xu[(size_t)NC * ip];
      double        cur[MAXS][MAXS][MAXS][4];
      double        sx0[MAXS], sx1[MAXS], sy0[MAXS], sy1[MAXS], sz0[MAXS], sz1[MAXS];
      int           bx, by = s->Lb[1], bz = s->Lb[0];
      const int     Sy = dim >= 2 ? S : 1, Sz = dim >= 3 ? S : 1;
      for (int jz = 0; jz < Sz; jz++)
        for (int jy = 0; jy < Sy; jy++)
          for (int jx = 0; jx < S; jx++)
            cur[jz][jy][jx][0] = cur[jz][jy][jx][1] = cur[jz][jy][jx][2] = cur[jz][jy][jx][3] = 0;

      bx = esirkepov_axis(s, 2, xv[0], xu[0], c->lim[2][0], sx0, sx1);
      if (dim == 1) {
        /* deposit1d, esirkepov.hpp:18-73 */
        const double vy = (xu[1] - xv[1]) / delt, vz = (xu[2] - xv[2]) / delt;
        for (int jx = 0; jx < S; jx++)
          cur[0][0][jx][0] += q * sx1[jx];
        for (int jx = 0; jx < S; jx++)
          sx1[jx] -= sx0[jx];
        const double qdxdt = q * (s->del[2] / delt), qvy = q * vy, qvz = q * vz;
        double       ww = 0, wx = -qdxdt;
        for (int jx = 0; jx < S - 1; jx++) {
          ww += sx1[jx] * wx;
          cur[0][0][jx + 1][1] += ww;
        }
        for (int jx = 0; jx < S; jx++)
          cur[0][0][jx][2] += (sx0[jx] + A * sx1[jx]) * qvy;
        for (int jx = 0; jx < S; jx++)
          cur[0][0][jx][3] += (sx0[jx] + A * sx1[jx]) * qvz;
      } else if (dim == 2) {
        /* deposit2d, esirkepov.hpp:76-150 */
        by = esirkepov_axis(s, 1, xv[1], xu[1], c->lim[1][0], sy0, sy1);
        const double vz = (xu[2] - xv[2]) / delt;
        for (int jy = 0; jy < S; jy++)
          for (int jx = 0; jx < S; jx++)
            cur[0][jy][jx][0] += q * sx1[jx] * sy1[jy];
        for (int j = 0; j < S; j++) {
          sx1[j] -= sx0[j];
          sy1[j] -= sy0[j];
        }
        const double qdxdt = q * (s->del[2] / delt), qdydt = q * (s->del[1] / delt), qvz = q * vz;
        for (int jy = 0; jy < S; jy++) {
          double ww = 0, wx = -(sy0[jy] + A * sy1[jy]) * qdxdt;
          for (int jx = 0; jx < S - 1; jx++) {
            ww += sx1[jx] * wx;
            cur[0][jy][jx + 1][1] += ww;
          }
        }
        for (int jx = 0; jx < S; jx++) {
          double ww = 0, wy = -(sx0[jx] + A * sx1[jx]) * qdydt;
          for (int jy = 0; jy < S - 1; jy++) {
            ww += sy1[jy] * wy;
            cur[0][jy + 1][jx][2] += ww;
          }
        }
        for (int jy = 0; jy < S; jy++)
          for (int jx = 0; jx < S; jx++)
            cur[0][jy][jx][3] +=
                ((1 * sx0[jx] + A * sx1[jx]) * sy0[jy] + (A * sx0[jx] + B * sx1[jx]) * sy1[jy]) * qvz;
      } else {
        /* deposit3d, esirkepov.hpp:152-237, 325-340 */
        by = esirkepov_axis(s, 1, xv[1], xu[1], c->lim[1][0], sy0, sy1);
        bz = esirkepov_axis(s, 0, xv[2], xu[2], c->lim[0][0], sz0, sz1);
        for (int jz = 0; jz < S; jz++)
          for (int jy = 0; jy < S; jy++)
            for (int jx = 0; jx < S; jx++)
              cur[jz][jy][jx][0] += q * sx1[jx] * sy1[jy] * sz1[jz];
        for (int j = 0; j < S; j++) {
          sx1[j] -= sx0[j];
          sy1[j] -= sy0[j];
          sz1[j] -= sz0[j];
        }
        const double qdxdt = q * (s->del[2] / delt), qdydt = q * (s->del[1] / delt),
                     qdzdt = q * (s->del[0] / delt);
        for (int jz = 0; jz < S; jz++)
          for (int jy = 0; jy < S; jy++) {
            double ww = 0;
            double wx = -((1 * sy0[jy] + A * sy1[jy]) * sz0[jz] + (A * sy0[jy] + B * sy1[jy]) * sz1[jz]) * qdxdt;
            for (int jx = 0; jx < S - 1; jx++) {
              ww += sx1[jx] * wx;
              cur[jz][jy][jx + 1][1] += ww;
            }
          }
        for (int jz = 0; jz < S; jz++)
          for (int jx = 0; jx < S; jx++) {
            double ww = 0;
            double wy = -((1 * sz0[jz] + A * sz1[jz]) * sx0[jx] + (A * sz0[jz] + B * sz1[jz]) * sx1[jx]) * qdydt;
            for (int jy = 0; jy < S - 1; jy++) {
              ww += sy1[jy] * wy;
              cur[jz][jy + 1][jx][2] += ww;
            }
          }
        for (int jy = 0; jy < S; jy++)
          for (int jx = 0; jx < S; jx++) {
            double ww = 0;
            double wz = -((1 * sx0[jx] + A * sx1[jx]) * sy0[jy] + (A * sx0[jx] + B * sx1[jx]) * sy1[jy]) * qdzdt;
            for (int jz = 0; jz < S - 1; jz++) {
              ww += sz1[jz] * wz;
              cur[jz + 1][jy][jx][3] += ww;
            }
          }
      }
      /* append_current{1,2,3}d, scalar branch, nix/primitives.hpp:678-834 */
      for (int jz = 0; jz < Sz; jz++)
        for (int jy = 0; jy < Sy; jy++)
          for (int jx = 0; jx < S; jx++)
            for (int k = 0; k < 4; k++)
              UJ(c, s, bz + jz, by + jy, bx + jx, k) += cur[jz][jy][jx][k];
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* moments: pic/engine/moment.hpp (scalar path), nix/primitives.hpp:836-930                   */
/* ------------------------------------------------------------------------------------------ */
static void chunk_deposit_moment(const orc_sim_t* s, chunk_t* c)
{
  const int    N = s->order + 1, Ns = s->cfg.Ns, dim = s->dimension;
  const double cc = s->cfg.cc, rc = 1 / cc;
  memset(c->um, 0, sizeof(double) * (size_t)s->Ng * Ns * 14);
  for (int is = 0; is < Ns; is++) {
    species_t* sp = &c->sp[is];
    for (int ip = 0; ip < sp->np; ip++) {
      const double* xu = &sp->xu[(size_t)NC * ip];
      double        wx[5], wy[5] = {1, 0, 0, 0, 0}, wz[5] = {1, 0, 0, 0, 0};
      int           ix0, iy0 = s->Lb[1], iz0 = s->Lb[0];
      {
        const double dx = s->del[2], rdx = 1 / dx, xmin = c->lim[2][0];
        ix0 = digitize(xu[0], xmin + 0.5 * dx * s->is_odd, rdx);
        shape_mc(s->order, xu[0], xmin + 0.5 * dx + ix0 * dx, rdx, wx);
        ix0 += s->Lb[2] - s->order / 2;
      }
      if (dim >= 2) {
        const double dx = s->del[1], rdx = 1 / dx, xmin = c->lim[1][0];
        iy0 = digitize(xu[1], xmin + 0.5 * dx * s->is_odd, rdx);
        shape_mc(s->order, xu[1], xmin + 0.5 * dx + iy0 * dx, rdx, wy);
        iy0 += s->Lb[1] - s->order / 2;
      }
      if (dim >= 3) {
        const double dx = s->del[0], rdx = 1 / dx, xmin = c->lim[0][0];
        iz0 = digitize(xu[2], xmin + 0.5 * dx * s->is_odd, rdx);
        shape_mc(s->order, xu[2], xmin + 0.5 * dx + iz0 * dx, rdx, wz);
        iz0 += s->Lb[0] - s->order / 2;
      }
      /* BaseMoment::local{1,2,3}d, pic/engine/moment.hpp:219-395; component indices :19-32 */
      const double ms = sp->m;
      const double gm = sqrt(1 + (xu[3] * xu[3] + xu[4] * xu[4] + xu[5] * xu[5]) * rc * rc);
      const int    Ny = dim >= 2 ? N : 1, Nz = dim >= 3 ? N : 1;
      for (int jz = 0; jz < Nz; jz++)
        for (int jy = 0; jy < Ny; jy++)
          for (int jx = 0; jx < N; jx++) {
            double ww = ms * wx[jx];
            if (dim >= 2)
              ww = ww * wy[jy];
            if (dim >= 3)
              ww = ww * wz[jz];
            double* m = &c->um[(CELL(s, iz0 + jz, iy0 + jy, ix0 + jx) * Ns + is) * 14];
            m[0] += ww;
            m[1] += ww * xu[3] / gm;
            m[2] += ww * xu[4] / gm;
            m[3] += ww * xu[5] / gm;
            m[8] += ww * xu[3];
            m[9] += ww * xu[4];
            m[10] += ww * xu[5];
            m[4] += ww * gm * cc;
            m[5] += ww * xu[3] * xu[3] / gm;
            m[6] += ww * xu[4] * xu[4] / gm;
            m[7] += ww * xu[5] * xu[5] / gm;
            m[11] += ww * xu[3] * xu[4] / gm;
            m[12] += ww * xu[4] * xu[5] / gm;
            m[13] += ww * xu[5] * xu[3] / gm;
          }
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* boundary exchange: nix/chunk.hpp:392-543, nix/xtensor_halo3d.hpp                           */
/* ------------------------------------------------------------------------------------------ */

static void dir_decode(int d, int dc[3])
{
  dc[0] = d / 9;
  dc[1] = (d / 3) % 3;
  dc[2] = d % 3;
}

/* copy/add a region of `src` chunk (starting at slo) into region of `dst` chunk (starting at dlo) */
static void region_apply(const orc_sim_t* s, double* dst, const int dlo[3], const double* src, const int slo[3],
                         const int len[3], int ncomp, int add)
{
  for (int jz = 0; jz < len[0]; jz++)
    for (int jy = 0; jy < len[1]; jy++)
      for (int jx = 0; jx < len[2]; jx++) {
        size_t cd = CELL(s, dlo[0] + jz, dlo[1] + jy, dlo[2] + jx) * ncomp;
        size_t cs = CELL(s, slo[0] + jz, slo[1] + jy, slo[2] + jx) * ncomp;
        for (int k = 0; k < ncomp; k++) {
          if (add)
            dst[cd + k] += src[cs + k];
          else
            dst[cd + k] = src[cs + k];
        }
      }
}

static void region_pack(const orc_sim_t* s, double* msg, const double* src, const int slo[3], const int len[3],
                        int ncomp)
{
  size_t e = 0;
  for (int jz = 0; jz < len[0]; jz++)
    for (int jy = 0; jy < len[1]; jy++)
      for (int jx = 0; jx < len[2]; jx++) {
        size_t cs = CELL(s, slo[0] + jz, slo[1] + jy, slo[2] + jx) * ncomp;
        for (int k = 0; k < ncomp; k++)
          msg[e++] = src[cs + k];
      }
}

static void region_unpack(const orc_sim_t* s, double* dst, const int dlo[3], const int len[3], const double* msg,
                          int ncomp, int add)
{
  size_t e = 0;
  for (int jz = 0; jz < len[0]; jz++)
    for (int jy = 0; jy < len[1]; jy++)
      for (int jx = 0; jx < len[2]; jx++) {
        size_t cd = CELL(s, dlo[0] + jz, dlo[1] + jy, dlo[2] + jx) * ncomp;
        for (int k = 0; k < ncomp; k++) {
          if (add)
            dst[cd + k] += msg[e++];
          else
            dst[cd + k] = msg[e++];
        }
      }
}

/* EMF copies interior margin -> ghost; CUR and MOM add ghost -> interior margin
 * (XtensorHaloField3D / Current3D / Moment3D, nix/xtensor_halo3d.hpp:18-185) */
static double* chunk_field(chunk_t* c, int mode)
{
  return mode == MODE_EMF ? c->uf : (mode == MODE_CUR ? c->uj : c->um);
}
static int mode_ncomp(const orc_sim_t* s, int mode)
{
  return mode == MODE_EMF ? 6 : (mode == MODE_CUR ? 4 : s->cfg.Ns * 14);
}

/* field / current: pack the messages for remote neighbours */
static void field_begin(orc_sim_t* s, int mode)
{
  const int ncomp = mode_ncomp(s, mode);
  for (int pi = 0; pi < s->npeer; pi++) {
    peer_t* p = &s->peers[pi];
    for (int m = 0; m < p->nsend; m++) {
      chunk_t* c = &s->chunks[p->send_chunk[m]];
      int      dc[3], lo[3], len[3];
      dir_decode(p->send_dir[m], dc);
      for (int a = 0; a < 3; a++) {
        len[a] = region_len(s, a, dc[a]);
        /* EMF sends the interior margin, CUR sends the ghost region (xtensor_halo3d.hpp:18-129) */
        lo[a] = mode == MODE_EMF ? margin_lo(s, a, dc[a]) : ghost_lo(s, a, dc[a]);
      }
      region_pack(s, p->send[mode] + p->send_off[mode][m], chunk_field(c, mode), lo, len, ncomp);
    }
  }
}

static int find_recv_msg(const peer_t* p, int ic, int d)
{
  for (int m = 0; m < p->nrecv; m++)
    if (p->recv_chunk[m] == ic && p->recv_dir[m] == d)
      return m;
  return -1;
}

/* unpack in the reference's order: directions ascending (nix/chunk.hpp:437-455) */
static void field_end(orc_sim_t* s, int mode)
{
  const int ncomp = mode_ncomp(s, mode);
  const int add   = mode != MODE_EMF;
#pragma omp parallel for schedule(dynamic) num_threads(s->nthread)
  for (int ic = 0; ic < s->nchunk; ic++) {
    chunk_t* c = &s->chunks[ic];
    for (int d = 0; d < 27; d++) {
      int dc[3], dlo[3], slo[3], len[3];
      dir_decode(d, dc);
      if (!dir_active(s, dc[0], dc[1], dc[2]) || c->nbid[d] < 0)
        continue;
      for (int a = 0; a < 3; a++) {
        len[a] = region_len(s, a, dc[a]);
        /* EMF: my ghost <- neighbour's interior margin on its opposite side;
         * CUR: my interior margin += neighbour's ghost on its opposite side */
        dlo[a] = mode == MODE_EMF ? ghost_lo(s, a, dc[a]) : margin_lo(s, a, dc[a]);
        slo[a] = mode == MODE_EMF ? margin_lo(s, a, 2 - dc[a]) : ghost_lo(s, a, 2 - dc[a]);
      }
      if (c->nbrank[d] == s->cfg.rank) {
        chunk_t* nbc = &s->chunks[c->nbid[d] - s->chunk_begin];
        region_apply(s, chunk_field(c, mode), dlo, chunk_field(nbc, mode), slo, len, ncomp, add);
      } else {
        peer_t* p = &s->peers[s->peer_of_rank[c->nbrank[d]]];
        int     m = find_recv_msg(p, ic, d);
        region_unpack(s, chunk_field(c, mode), dlo, len, p->recv[mode] + p->recv_off[mode][m], ncomp, add);
      }
    }
  }
}

/* particle records on the wire: 7 components + tag {destination chunk id, species | sender dir << 8} */
static void particle_begin(orc_sim_t* s)
{
  const int Ns = s->cfg.Ns;
  /* XtensorHaloParticle3D::pre_pack: direction code per particle, xtensor_halo3d.hpp:214-348 */
#pragma omp parallel for schedule(dynamic) num_threads(s->nthread)
  for (int ic = 0; ic < s->nchunk; ic++) {
    chunk_t* c = &s->chunks[ic];
    for (int d = 0; d < 27; d++) {
      memset(c->out_cnt[d], 0, sizeof(int) * Ns);
    }
    int total[27] = {0};
    for (int pass = 0; pass < 2; pass++) {
      int fill[27] = {0};
      for (int is = 0; is < Ns; is++) {
        species_t* sp = &c->sp[is];
        for (int ip = 0; ip < sp->np; ip++) {
          const double* p  = &sp->xu[(size_t)NC * ip];
          int           ix = s->has_dim[2] ? (p[0] >= c->lim[2][1]) - (p[0] < c->lim[2][0]) + 1 : 1;
          int           iy = s->has_dim[1] ? (p[1] >= c->lim[1][1]) - (p[1] < c->lim[1][0]) + 1 : 1;
          int           iz = s->has_dim[0] ? (p[2] >= c->lim[0][1]) - (p[2] < c->lim[0][0]) + 1 : 1;
          int           d  = 9 * iz + 3 * iy + ix;
          if (d == 13)
            continue;
          if (pass == 0) {
            total[d]++;
            c->out_cnt[d][is]++;
          } else {
            c->out_idx[d][fill[d]++] = ip;
          }
        }
      }
      if (pass == 0) {
        for (int d = 0; d < 27; d++)
          if (total[d] > c->out_cap[d]) {
            c->out_cap[d] = total[d] + 64;
            c->out_idx[d] = (int*)realloc(c->out_idx[d], sizeof(int) * c->out_cap[d]);
          }
      }
    }
  }
  /* remote: serialise per peer in (chunk, dir, species, ip) order */
  for (int pi = 0; pi < s->npeer; pi++) {
    peer_t* p   = &s->peers[pi];
    int64_t nrec = 0;
    for (int m = 0; m < p->nsend; m++) {
      chunk_t* c = &s->chunks[p->send_chunk[m]];
      for (int is = 0; is < Ns; is++)
        nrec += c->out_cnt[p->send_dir[m]][is];
    }
    if (nrec > p->psend_cap) {
      p->psend_cap = nrec + 1024;
      p->psend     = (double*)realloc(p->psend, sizeof(double) * 8 * p->psend_cap);
    }
    int64_t r = 0;
    for (int m = 0; m < p->nsend; m++) {
      chunk_t* c   = &s->chunks[p->send_chunk[m]];
      int      d   = p->send_dir[m];
      int      pos = 0;
      for (int is = 0; is < Ns; is++) {
        for (int k = 0; k < c->out_cnt[d][is]; k++) {
          int     ip  = c->out_idx[d][pos++];
          double* out = p->psend + 8 * r++;
          memcpy(out, &c->sp[is].xu[(size_t)NC * ip], sizeof(double) * NC);
          int32_t tag[2] = {c->nbid[d], is | (d << 8)};
          memcpy(out + 7, tag, 8);
        }
      }
    }
    p->psend_bytes = r * 64;
    p->precv_bytes = 0;
  }
}

static void particle_end(orc_sim_t* s)
{
  const int Ns = s->cfg.Ns;
  /* bucket the received records by (local chunk, receive direction, species), keeping their order */
  int     nkey  = s->nchunk * 27 * Ns;
  int*    rcnt  = (int*)calloc(nkey + 1, sizeof(int));
  int*    roff  = (int*)calloc(nkey + 1, sizeof(int));
  int64_t total = 0;
  for (int pi = 0; pi < s->npeer; pi++)
    total += s->peers[pi].precv_bytes / 64;
  const double** rptr = (const double**)malloc(sizeof(double*) * (total + 1));
  for (int pass = 0; pass < 2; pass++) {
    for (int pi = 0; pi < s->npeer; pi++) {
      peer_t* p = &s->peers[pi];
      int64_t n = p->precv_bytes / 64;
      for (int64_t r = 0; r < n; r++) {
        const double* rec = p->precv + 8 * r;
        int32_t       tag[2];
        memcpy(tag, rec + 7, 8);
        int ic = tag[0] - s->chunk_begin, is = tag[1] & 0xff, sd = tag[1] >> 8;
        int key = (ic * 27 + (26 - sd)) * Ns + is;
        if (pass == 0)
          rcnt[key]++;
        else
          rptr[roff[key]++] = rec;
      }
    }
    if (pass == 0) {
      int acc = 0;
      for (int k = 0; k < nkey; k++) {
        roff[k] = acc;
        acc += rcnt[k];
      }
    } else {
      int acc = 0;
      for (int k = 0; k < nkey; k++) {
        roff[k] = acc;
        acc += rcnt[k];
      }
    }
  }

  /* unpack: directions ascending, species inside a message, append behind Np (xtensor_halo3d.hpp:426-475) */
#pragma omp parallel for schedule(dynamic) num_threads(s->nthread)
  for (int ic = 0; ic < s->nchunk; ic++) {
    chunk_t* c = &s->chunks[ic];
    for (int is = 0; is < Ns; is++) {
      int incoming = 0;
      for (int d = 0; d < 27; d++) {
        int dc[3];
        dir_decode(d, dc);
        if (!dir_active(s, dc[0], dc[1], dc[2]) || c->nbid[d] < 0)
          continue;
        if (c->nbrank[d] == s->cfg.rank)
          incoming += s->chunks[c->nbid[d] - s->chunk_begin].out_cnt[26 - d][is];
        else
          incoming += rcnt[(ic * 27 + d) * Ns + is];
      }
      c->sp[is].ntail = incoming;
    }
  }
  /* growing a buffer moves it, so no other thread may be reading: serial */
  for (int ic = 0; ic < s->nchunk; ic++)
    for (int is = 0; is < Ns; is++) {
      species_t* sp = &s->chunks[ic].sp[is];
      species_reserve(sp, sp->np + sp->ntail); /* XtensorParticle::resize, :70-115 */
    }
#pragma omp parallel for schedule(dynamic) num_threads(s->nthread)
  for (int ic = 0; ic < s->nchunk; ic++) {
    chunk_t* c = &s->chunks[ic];
    int      unpacked[64] = {0};
    for (int d = 0; d < 27; d++) {
      int dc[3];
      dir_decode(d, dc);
      if (!dir_active(s, dc[0], dc[1], dc[2]) || c->nbid[d] < 0)
        continue;
      if (c->nbrank[d] == s->cfg.rank) {
        chunk_t* nbc = &s->chunks[c->nbid[d] - s->chunk_begin];
        int      od = 26 - d, pos = 0;
        for (int is = 0; is < Ns; is++)
          for (int k = 0; k < nbc->out_cnt[od][is]; k++) {
            int ip = nbc->out_idx[od][pos++];
            memcpy(&c->sp[is].xu[(size_t)NC * (c->sp[is].np + unpacked[is]++)],
                   &nbc->sp[is].xu[(size_t)NC * ip], sizeof(double) * NC);
          }
      } else {
        for (int is = 0; is < Ns; is++) {
          int key = (ic * 27 + d) * Ns + is;
          for (int k = 0; k < rcnt[key]; k++)
            memcpy(&c->sp[is].xu[(size_t)NC * (c->sp[is].np + unpacked[is]++)], rptr[roff[key] + k],
                   sizeof(double) * NC);
        }
      }
    }
  }
  /* post_unpack: periodic wrap + count of the received particles, then sort (:477-498) */
#pragma omp parallel for schedule(dynamic) num_threads(s->nthread)
  for (int ic = 0; ic < s->nchunk; ic++) {
    chunk_t* c = &s->chunks[ic];
    for (int is = 0; is < Ns; is++) {
      species_t* sp      = &c->sp[is];
      int        np_prev = sp->np, np_next = sp->np + sp->ntail;
      /* set_boundary_periodic, nix/xtensor_particle.hpp:359-376 */
      const double X = s->has_dim[2] * (s->glim[2][1] - s->glim[2][0]);
      const double Y = s->has_dim[1] * (s->glim[1][1] - s->glim[1][0]);
      const double Z = s->has_dim[0] * (s->glim[0][1] - s->glim[0][0]);
      for (int ip = np_prev; ip < np_next; ip++) {
        double* p = &sp->xu[(size_t)NC * ip];
        p[0] += (p[0] < s->glim[2][0]) * X - (p[0] >= s->glim[2][1]) * X;
        p[1] += (p[1] < s->glim[1][0]) * Y - (p[1] >= s->glim[1][1]) * Y;
        p[2] += (p[2] < s->glim[0][0]) * Z - (p[2] >= s->glim[0][1]) * Z;
      }
      species_count(s, c, sp, np_prev, np_next - 1, 0);
      sp->np    = np_next;
      sp->ntail = 0;
    }
  }
#pragma omp parallel for schedule(dynamic) num_threads(s->nthread)
  for (int ic = 0; ic < s->nchunk; ic++)
    for (int is = 0; is < Ns; is++)
      species_sort(s, &s->chunks[ic].sp[is]);
  free(rcnt);
  free(roff);
  free(rptr);
}

void orc_boundary_begin(orc_sim_t* s, int mode)
{
  if (mode == MODE_EMF || mode == MODE_CUR || mode == MODE_MOM)
    field_begin(s, mode);
  else if (mode == MODE_PARTICLE)
    particle_begin(s);
}

void orc_boundary_end(orc_sim_t* s, int mode)
{
  if (mode == MODE_EMF || mode == MODE_CUR || mode == MODE_MOM)
    field_end(s, mode);
  else if (mode == MODE_PARTICLE)
    particle_end(s);
}

int orc_get_peers(const orc_sim_t* s, int32_t* peer_rank)
{
  if (peer_rank != NULL)
    for (int i = 0; i < s->npeer; i++)
      peer_rank[i] = s->peers[i].rank;
  return s->npeer;
}

void orc_get_comm_buffer(orc_sim_t* s, int mode, int peer_index, void** send_ptr, int64_t* send_bytes,
                         void** recv_ptr, int64_t* recv_bytes)
{
  peer_t* p = &s->peers[peer_index];
  if (mode == MODE_PARTICLE) {
    *send_ptr   = p->psend;
    *send_bytes = p->psend_bytes;
    *recv_ptr   = p->precv;
    *recv_bytes = p->precv_bytes;
  } else {
    *send_ptr   = p->send[mode];
    *send_bytes = p->send_elems[mode] * 8;
    *recv_ptr   = p->recv[mode];
    *recv_bytes = p->recv_elems[mode] * 8;
  }
}

void orc_set_recv_bytes(orc_sim_t* s, int mode, int peer_index, int64_t recv_bytes)
{
  peer_t* p = &s->peers[peer_index];
  if (mode != MODE_PARTICLE)
    return;
  int64_t nrec = recv_bytes / 64;
  if (nrec > p->precv_cap) {
    p->precv_cap = nrec + 1024;
    p->precv     = (double*)realloc(p->precv, sizeof(double) * 8 * p->precv_cap);
  }
  p->precv_bytes = recv_bytes;
}

/* ------------------------------------------------------------------------------------------ */
/* phases over all local chunks                                                               */
/* ------------------------------------------------------------------------------------------ */

#define FOR_CHUNKS(s, body)                                                                        \
  _Pragma("omp parallel for schedule(dynamic) num_threads(s->nthread)") for (int ic_ = 0; ic_ < (s)->nchunk; ic_++)  \
  {                                                                                                \
    chunk_t* c = &(s)->chunks[ic_];                                                                \
    body;                                                                                          \
  }

void orc_init_friedman(orc_sim_t* s) { FOR_CHUNKS(s, chunk_init_friedman(s, c)); }
void orc_push_bfd(orc_sim_t* s, double delt) { FOR_CHUNKS(s, chunk_push_bfd(s, c, delt)); }
void orc_push_efd(orc_sim_t* s, double delt) { FOR_CHUNKS(s, chunk_push_efd(s, c, delt)); }
void orc_push_velocity(orc_sim_t* s, double delt) { FOR_CHUNKS(s, chunk_push_velocity(s, c, delt)); }
void orc_push_position(orc_sim_t* s, double delt) { FOR_CHUNKS(s, chunk_push_position(s, c, delt)); }
void orc_deposit_current(orc_sim_t* s, double delt) { FOR_CHUNKS(s, chunk_deposit_current(s, c, delt)); }
void orc_deposit_moment(orc_sim_t* s) { FOR_CHUNKS(s, chunk_deposit_moment(s, c)); }

/* PicChunk::sort_particle, pic/pic_chunk.cpp:447-453 */
void orc_sort_particle(orc_sim_t* s)
{
  FOR_CHUNKS(s, {
    for (int is = 0; is < s->cfg.Ns; is++) {
      species_count(s, c, &c->sp[is], 0, c->sp[is].np - 1, 1);
      species_sort(s, &c->sp[is]);
    }
  });
}

void orc_exchange(orc_sim_t* s, int mode)
{
  orc_boundary_begin(s, mode);
  orc_boundary_end(s, mode);
}

/* PicApplication::push_openmp, pic/pic_application.cpp:219-292 (single rank: no transport needed) */
void orc_step(orc_sim_t* s, double delt, int nstep)
{
  for (int step = 0; step < nstep; step++) {
    orc_push_bfd(s, 0.5 * delt);
    orc_push_velocity(s, delt);
    orc_push_position(s, delt);
    orc_deposit_current(s, delt);
    orc_boundary_begin(s, MODE_CUR);
    orc_boundary_begin(s, MODE_PARTICLE);
    orc_push_bfd(s, 0.5 * delt);
    orc_boundary_end(s, MODE_CUR);
    orc_push_efd(s, delt);
    orc_boundary_begin(s, MODE_EMF);
    orc_boundary_end(s, MODE_PARTICLE);
    orc_boundary_end(s, MODE_EMF);
  }
}
