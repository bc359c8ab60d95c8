// -*- C++ -*-
// Dispatch of the tiled (row-owner) push + deposit kernels:
//   3-D, order 2 : rowpush.cu    (row_push_kernel)
//   2-D, order 2 : rowpush2d.cu  (row_push2d_kernel)
//   1-D, order 2 : rowpush1d.cu  (row_push1d_kernel)
// Everything else (other orders, rows that do not split into segments) runs the generic kernels of
// particle.cu / fused.cu.  experiments/rowfused_v1.cu holds the round-1 formulation and the FP64-MMA
// experiment; they are reachable through options only.
#include "arena.hpp"

namespace picnix
{

bool row_push2d_geometry(const picnix_arena* a); // rowpush2d.cu
int  launch_deposit_rows_2d(picnix_arena* a, int c0, int cn, double delt);
int  launch_row_fused_2d(picnix_arena* a, int c0, int cn, double delt);
bool row_push1d_geometry(const picnix_arena* a); // rowpush1d.cu
int  launch_deposit_rows_1d(picnix_arena* a, int c0, int cn, double delt);
int  launch_row_fused_1d(picnix_arena* a, int c0, int cn, double delt);
bool row_push_applies(const picnix_arena* a); // rowpush.cu: at most rowtile::MAXNS species
int  launch_deposit_rows_v2(picnix_arena* a, int c0, int cn, double delt);
int  launch_row_fused_v2(picnix_arena* a, int c0, int cn, double delt);
bool row_v1_geometry(const picnix_arena* a); // experiments/rowfused_v1.cu (same row geometry as rowpush.cu)
int  launch_row_kernel_v1(picnix_arena* a, int c0, int cn, double delt, bool fused);

// The tiled kernels need 2nd-order shapes, rows that split into 8-cell segments and 4-row groups, and a
// pindex that describes the current particle order (set by the sort, cleared by uploads).
bool row_geometry_applies(const picnix_arena* a)
{
  return row_push2d_geometry(a) || row_push1d_geometry(a) || row_v1_geometry(a);
}

bool row_kernel_applies(const picnix_arena* a)
{
  return row_geometry_applies(a) && a->pindex_valid && !a->force_generic;
}

int launch_deposit_rows(picnix_arena* a, int c0, int cn, double delt)
{
  if (a->g.dimension == 2)
    return launch_deposit_rows_2d(a, c0, cn, delt);
  if (a->g.dimension == 1)
    return launch_deposit_rows_1d(a, c0, cn, delt);
  if (a->row_version >= 2 && !a->deposit_mma && row_push_applies(a))
    return launch_deposit_rows_v2(a, c0, cn, delt);
  return launch_row_kernel_v1(a, c0, cn, delt, false);
}

int launch_row_fused(picnix_arena* a, int c0, int cn, double delt)
{
  if (a->g.dimension == 2)
    return launch_row_fused_2d(a, c0, cn, delt);
  if (a->g.dimension == 1)
    return launch_row_fused_1d(a, c0, cn, delt);
  // physical boundary conditions are handled by rowpush.cu only
  if ((a->row_version >= 2 || a->any_bc) && !a->deposit_mma && row_push_applies(a))
    return launch_row_fused_v2(a, c0, cn, delt);
  return launch_row_kernel_v1(a, c0, cn, delt, true);
}

} // namespace picnix
