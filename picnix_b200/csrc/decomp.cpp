// -*- C++ -*-
// Host-side domain decomposition: space-filling-curve chunk ordering and rank assignment.
//
// Integer logic only; results must be bit-exact with the reference:
//   * generalized Hilbert ("gilbert") curve      nix/sfc.cpp:31-141, 200-508
//   * initial / incremental rank boundaries      nix/balancer.cpp:8-124
// The curve is the published algorithm of J. Cerveny (generalized Hilbert curve for arbitrary
// rectangular domains); it is written here on small integer vectors rather than on nine scalar
// arguments.
#include "arena.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace picnix
{
namespace
{

struct V3 {
  int x, y, z;
};

inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 half(V3 a) { return {a.x / 2, a.y / 2, a.z / 2}; } // truncating, like the reference
inline int sgn(int v) { return v == 0 ? 0 : (v > 0 ? 1 : -1); }
inline V3 sgn(V3 a) { return {sgn(a.x), sgn(a.y), sgn(a.z)}; }
inline int len(V3 a) { return std::abs(a.x + a.y + a.z); }

// visits cells in curve order and numbers them
struct Walker {
  std::vector<int32_t>& index;
  int                   Ny, Nx;
  int                   id = 0;

  void visit(V3 p) { index[(size_t)(p.z * Ny + p.y) * Nx + p.x] = id++; }

  void line(V3 p, V3 step, int n)
  {
    for (int i = 0; i < n; i++) {
      visit(p);
      p = p + step;
    }
  }

  // two-dimensional curve spanned by major axis a and minor axis b (nix/sfc.cpp:200-294)
  void curve2(V3 p, V3 a, V3 b)
  {
    const int w = len(a), h = len(b);
    const V3  da = sgn(a), db = sgn(b);

    if (h == 1) {
      line(p, da, w);
      return;
    }
    if (w == 1) {
      line(p, db, h);
      return;
    }

    V3 a2 = half(a), b2 = half(b);

    if (2 * w > 3 * h) {
      // long in a: split along a only
      if ((len(a2) % 2) && (w > 2))
        a2 = a2 + da;
      curve2(p, a2, b);
      curve2(p + a2, a - a2, b);
    } else {
      if ((len(b2) % 2) && (h > 2))
        b2 = b2 + db;
      curve2(p, b2, a2);
      curve2(p + b2, a, b - b2);
      curve2(p + b2 + a - da - db, -b2, -(a - a2));
    }
  }

  // three-dimensional curve (nix/sfc.cpp:296-508)
  void curve3(V3 p, V3 a, V3 b, V3 c)
  {
    const int w = len(a), h = len(b), d = len(c);
    const V3  da = sgn(a), db = sgn(b), dc = sgn(c);

    if (h == 1 && d == 1) {
      line(p, da, w);
      return;
    }
    if (d == 1 && w == 1) {
      line(p, db, h);
      return;
    }
    if (w == 1 && h == 1) {
      line(p, dc, d);
      return;
    }

    V3 a2 = half(a), b2 = half(b), c2 = half(c);
    if ((len(a2) % 2) && (w > 2))
      a2 = a2 + da;
    if ((len(b2) % 2) && (h > 2))
      b2 = b2 + db;
    if ((len(c2) % 2) && (d > 2))
      c2 = c2 + dc;
    const V3 a3 = a - a2, b3 = b - b2, c3 = c - c2;

    if ((2 * w > 3 * h) && (2 * w > 3 * d)) {
      // split along a only
      curve3(p, a2, b, c);
      curve3(p + a2, a3, b, c);
    } else if (3 * h > 4 * d) {
      // split in the a-b plane
      curve3(p, b2, c, a2);
      p = p + b2;
      curve3(p, a, b3, c);
      p = p + a - da - db;
      curve3(p, -b2, c, -a3);
    } else if (3 * d > 4 * h) {
      // split in the a-c plane
      curve3(p, c2, a2, b);
      p = p + c2;
      curve3(p, a, b, c3);
      p = p + a - da - dc;
      curve3(p, -c2, -a3, b);
    } else {
      // full three-dimensional split
      curve3(p, b2, c2, a2);
      p = p + b2;
      curve3(p, c, a2, b3);
      p = p + c - db - dc;
      curve3(p, a, -b2, -c3);
      p = p + a - (da - db);
      curve3(p, -c, -a3, b3);
      p = p - c - (db - dc);
      curve3(p, -b2, c2, -a3);
    }
  }
};

} // namespace

void sfc_build(int Cz, int Cy, int Cx, std::vector<int32_t>& chunkid, std::vector<int32_t>& coord)
{
  const size_t n = (size_t)Cz * Cy * Cx;
  chunkid.assign(n, 0);
  coord.assign(3 * n, 0);

  Walker   walk{chunkid, Cy, Cx};
  const V3 origin{0, 0, 0};
  const V3 ex{Cx, 0, 0}, ey{0, Cy, 0}, ez{0, 0, Cz};
  const int nlong = (Cx != 1) + (Cy != 1) + (Cz != 1);

  if (nlong == 3) {
    // the longest extent leads (nix/sfc.cpp:88-94)
    if (Cx >= Cy && Cx >= Cz) {
      walk.curve3(origin, ex, ey, ez);
    } else if (Cy >= Cx && Cy >= Cz) {
      walk.curve3(origin, ey, ex, ez);
    } else {
      walk.curve3(origin, ez, ex, ey);
    }
  } else if (nlong == 2) {
    // planar curve in the two non-degenerate directions (nix/sfc.cpp:39-52, 110-124):
    // "u" is the faster-varying of the two, "v" the slower one
    V3 eu, ev;
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
    if (len(eu) >= len(ev)) {
      walk.curve2(origin, eu, ev);
    } else {
      walk.curve2(origin, ev, eu);
    }
  } else if (nlong == 1) {
    // plain ordering along the only direction (nix/sfc.cpp:31-37)
    for (size_t i = 0; i < n; i++)
      chunkid[i] = (int32_t)i;
  } else {
    chunkid[0] = 0;
  }

  // id -> coordinate, stored (x, y, z) like nix::ChunkMap::coord (nix/sfc.cpp:130-140)
  for (int iz = 0; iz < Cz; iz++) {
    for (int iy = 0; iy < Cy; iy++) {
      for (int ix = 0; ix < Cx; ix++) {
        int id            = chunkid[(size_t)(iz * Cy + iy) * Cx + ix];
        coord[3 * id + 0] = ix;
        coord[3 * id + 1] = iy;
        coord[3 * id + 2] = iz;
      }
    }
  }
}

//
// rank boundaries
//

static std::vector<double> cumulative(const std::vector<double>& load)
{
  std::vector<double> cum(load.size() + 1);
  cum[0] = 0;
  for (size_t i = 0; i < load.size(); i++)
    cum[i + 1] = cum[i] + load[i];
  return cum;
}

// nix/balancer.cpp:71-99 ; returns false when the boundaries are not strictly ascending
bool assign_binarysearch(const std::vector<double>& load, std::vector<int32_t>& boundary)
{
  const int nc  = (int)load.size();
  const int nr  = (int)boundary.size() - 1;
  auto      cum = cumulative(load);
  double    mean = cum[nc] / nr;

  boundary[0]  = 0;
  boundary[nr] = nc;
  for (int i = 1; i < nr; i++) {
    auto it     = std::upper_bound(cum.begin(), cum.end(), mean * i);
    boundary[i] = (int32_t)(it - cum.begin()) - 1;
  }

  // Balancer::is_boundary_ascending (nix/balancer.cpp:166-180) starts its check at i = 1
  bool ascending = (boundary[0] == 0) && (boundary[nr] == nc);
  for (int i = 1; i < nr; i++) {
    ascending = ascending && (boundary[i + 1] > boundary[i]);
  }
  return ascending;
}

// nix/balancer.cpp:8-69 ; returns true when a boundary moved
bool assign_smilei(const std::vector<double>& load, std::vector<int32_t>& boundary)
{
  const int            nc  = (int)load.size();
  const int            nr  = (int)boundary.size() - 1;
  auto                 cum = cumulative(load);
  double               mean = cum[nc] / nr;
  std::vector<int32_t> old(boundary);

  for (int i = 1; i < nr; i++) {
    double target  = mean * i;
    double current = cum[boundary[i]];

    if (current > target) {
      // try to pull the boundary back
      int index = boundary[i] - 1;
      while (std::abs(current - target) > std::abs(current - target - load[index])) {
        current -= load[index];
        index--;
      }
      boundary[i] = (index >= old[i - 1]) ? index + 1 : old[i - 1] + 1;
    } else {
      // push the boundary forward
      int index = boundary[i];
      while (std::abs(current - target) > std::abs(current - target + load[index])) {
        current += load[index];
        index++;
      }
      boundary[i] = (index < old[i + 1]) ? index : old[i + 1] - 1;
    }
  }

  return !std::equal(boundary.begin(), boundary.end(), old.begin());
}

// nix/balancer.cpp:101-124
std::vector<int32_t> assign_initial(const std::vector<double>& load, int nrank)
{
  std::vector<int32_t> boundary(nrank + 1);

  if (!assign_binarysearch(load, boundary)) {
    std::vector<double> uniform(load.size(), 1.0);
    assign_binarysearch(uniform, boundary);
    for (int iter = 0; iter < 100; iter++) {
      if (!assign_smilei(load, boundary))
        break;
    }
  }
  return boundary;
}

} // namespace picnix

extern "C" {

int picnix_sfc_build(int32_t Cz, int32_t Cy, int32_t Cx, int32_t* chunkid, int32_t* coord)
{
  if (Cz < 1 || Cy < 1 || Cx < 1 || chunkid == nullptr || coord == nullptr)
    return PICNIX_ERR_INVALID;
  std::vector<int32_t> id, co;
  picnix::sfc_build(Cz, Cy, Cx, id, co);
  std::copy(id.begin(), id.end(), chunkid);
  std::copy(co.begin(), co.end(), coord);
  return PICNIX_OK;
}

int picnix_assign_initial(const double* load, int32_t nchunk, int32_t nrank, int32_t* boundary)
{
  if (load == nullptr || boundary == nullptr || nchunk < 1 || nrank < 1 || nrank > nchunk)
    return PICNIX_ERR_INVALID;
  std::vector<double> l(load, load + nchunk);
  auto                b = picnix::assign_initial(l, nrank);
  std::copy(b.begin(), b.end(), boundary);
  return PICNIX_OK;
}

int picnix_assign_rebalance(const double* load, int32_t nchunk, int32_t nrank, int32_t* boundary)
{
  if (load == nullptr || boundary == nullptr || nchunk < 1 || nrank < 1 || nrank > nchunk)
    return PICNIX_ERR_INVALID;
  std::vector<double>  l(load, load + nchunk);
  std::vector<int32_t> b(boundary, boundary + nrank + 1);
  picnix::assign_smilei(l, b);
  std::copy(b.begin(), b.end(), boundary);
  return PICNIX_OK;
}

} // extern "C"
