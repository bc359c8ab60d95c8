// -*- C++ -*-
// Physical boundary conditions on the faces of a non-periodic global domain, on the device.
//
// In the reference these are problem code: virtual hooks of the example's MainChunk that PicChunk calls
// after every halo unpack (set_boundary_field, pic/pic_chunk.cpp:360-361) and after the position push
// (set_boundary_particle, pic/pic_engine.hpp:292-303).  The two examples of BASELINE.json that are not
// periodic use three kinds, provided here as built-in kinds a problem selects per (axis, side) through
// picnix_cuda_set_boundary_condition:
//
//   PICNIX_BC_CONDUCTING  example/mrx/main.cpp:183-300 -- conducting wall normal to y: tangential E
//                         antisymmetric and tangential B symmetric about the mirror plane, normal E
//                         from Gauss' law with the deposited charge, normal B from div B = 0; particles
//                         are reflected specularly (:352-382)
//   PICNIX_BC_WALL        example/shock/main.cpp:232-283 -- wall at the lower x boundary: E = 0 and B
//                         continued into the margin; particles bounce back with all momentum components
//                         reversed (:419-433)
//   PICNIX_BC_INFLOW      example/shock/main.cpp:285-340 -- upstream values imposed in the margin of the
//                         upper x boundary; particles leaving there are dropped by the sort (the
//                         re-injection is host code with the host's random numbers: picnix_cuda_inject_particles)
//
// The index ranges below are the reference's, including its asymmetries between the lower and the upper
// side (the mirror plane lies `margin` cells inside the domain, the stagger of the normal components).
// The particle part is apply_particle_bc() in particle_common.cuh, called by every position push.
// The moment margins (BoundaryMom) are left as the halo exchange produced them: example/mrx indexes the
// five-dimensional moment array with four indices there, which is not reproduced.
#include "particle_common.cuh"

namespace picnix
{

namespace
{

constexpr int BC_THREADS = 128;

__device__ __forceinline__ double& F(double* uf, const Geom& g, int iz, int iy, int ix, int k)
{
  return uf[((int64_t)(iz * g.M[1] + iy) * g.M[2] + ix) * 6 + k];
}

// ---- conducting wall normal to y ---------------------------------------------------------------
// pass 0: tangential components, independent per cell; pass 1: normal components, sequential along y
// in every (z, x) column and reading the tangential values of pass 0 of the neighbouring columns
__global__ void __launch_bounds__(BC_THREADS)
conducting_y_kernel(Geom g, DevPtrs d, int side, int pass)
{
  const int chunk = blockIdx.y;
  const int dir   = 9 * 1 + 3 * (side == 0 ? 0 : 2) + 1; // neighbour in -y / +y
  if (d.nbr[chunk * NBSIZE + dir] != NB_NONE)
    return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.M[0] * g.M[2])
    return;
  const int iz = t / g.M[2], ix = t - iz * g.M[2]; // whole padded extent: Lb - Nb .. Ub + Nb
  const int Nb = g.nb, Lby = g.Lb[1], Uby = g.Ub[1];
  double*       uf = d.uf + (int64_t)chunk * g.Ng * 6;
  const double* uj = d.uj + (int64_t)chunk * g.Ng * 4;
  const double  delyx = g.del[1] / g.del[2] * g.has_dim[2];
  const double  delyz = g.del[1] / g.del[0] * g.has_dim[0];
  const double  dely  = g.del[1];
  auto rho = [&](int jz, int jy, int jx) { return uj[((int64_t)(jz * g.M[1] + jy) * g.M[2] + jx) * 4 + 0]; };

  if (side == 0) {
    if (pass == 0) {
      for (int iy = 0; iy < 2 * Nb; iy++) {
        const int iy1 = Lby - iy + Nb - 1;
        F(uf, g, iz, iy1, ix, 0) = -F(uf, g, iz, Lby + iy + Nb, ix, 0);
        F(uf, g, iz, iy1, ix, 2) = -F(uf, g, iz, Lby + iy + Nb, ix, 2);
        F(uf, g, iz, iy1, ix, 3) = F(uf, g, iz, Lby + iy + Nb + 1, ix, 3);
        F(uf, g, iz, iy1, ix, 5) = F(uf, g, iz, Lby + iy + Nb + 1, ix, 5);
      }
    } else {
      if (iz <= g.M[0] - 2 && ix <= g.M[2] - 2) {
        for (int iy = 0; iy < 2 * Nb; iy++) {
          const int iy1 = Lby - iy + Nb - 1, iy2 = iy1 + 1;
          F(uf, g, iz, iy1, ix, 1) = -dely * rho(iz, iy1, ix) + F(uf, g, iz, iy2, ix, 1) +
                                     delyx * (F(uf, g, iz, iy1, ix + 1, 0) - F(uf, g, iz, iy1, ix, 0)) +
                                     delyz * (F(uf, g, iz + 1, iy1, ix, 2) - F(uf, g, iz, iy1, ix, 2));
        }
      }
      if (iz >= 1 && ix >= 1) {
        for (int iy = 0; iy < 2 * Nb; iy++) {
          const int iy1 = Lby - iy + Nb - 1, iy2 = iy1 + 1;
          F(uf, g, iz, iy1, ix, 4) = F(uf, g, iz, iy2, ix, 4) +
                                     delyx * (F(uf, g, iz, iy2, ix, 3) - F(uf, g, iz, iy2, ix - 1, 3)) +
                                     delyz * (F(uf, g, iz, iy2, ix, 5) - F(uf, g, iz - 1, iy2, ix, 5));
        }
      }
    }
  } else {
    if (pass == 0) {
      for (int iy = 0; iy < 2 * Nb; iy++) {
        const int iy1 = Uby + iy - Nb + 1, iy2 = Uby - iy - Nb;
        F(uf, g, iz, iy1, ix, 0) = -F(uf, g, iz, iy2, ix, 0);
        F(uf, g, iz, iy1, ix, 2) = -F(uf, g, iz, iy2, ix, 2);
      }
      for (int iy = 0; iy < 2 * Nb - 1; iy++) {
        const int iy1 = Uby + iy - Nb + 2, iy2 = Uby - iy - Nb;
        F(uf, g, iz, iy1, ix, 3) = F(uf, g, iz, iy2, ix, 3);
        F(uf, g, iz, iy1, ix, 5) = F(uf, g, iz, iy2, ix, 5);
      }
    } else {
      if (iz <= g.M[0] - 2 && ix <= g.M[2] - 2) {
        for (int iy = 0; iy < 2 * Nb - 1; iy++) {
          const int iy1 = Uby + iy - Nb + 2, iy2 = iy1 - 1;
          F(uf, g, iz, iy1, ix, 1) = +dely * rho(iz, iy2, ix) + F(uf, g, iz, iy2, ix, 1) -
                                     delyx * (F(uf, g, iz, iy2, ix + 1, 0) - F(uf, g, iz, iy2, ix, 0)) -
                                     delyz * (F(uf, g, iz + 1, iy2, ix, 2) - F(uf, g, iz, iy2, ix, 2));
        }
      }
      if (iz >= 1 && ix >= 1) {
        for (int iy = 0; iy < 2 * Nb; iy++) {
          const int iy1 = Uby + iy - Nb + 1, iy2 = iy1 - 1;
          F(uf, g, iz, iy1, ix, 4) = F(uf, g, iz, iy2, ix, 4) -
                                     delyx * (F(uf, g, iz, iy1, ix, 3) - F(uf, g, iz, iy1, ix - 1, 3)) -
                                     delyz * (F(uf, g, iz, iy1, ix, 5) - F(uf, g, iz - 1, iy1, ix, 5));
        }
      }
    }
  }
}

// ---- wall (lower x) and inflow (upper x) of the shock tube ---------------------------------------
__global__ void __launch_bounds__(BC_THREADS)
shock_x_kernel(Geom g, DevPtrs d, int side, int kind)
{
  const int chunk = blockIdx.y;
  const int dir   = 9 * 1 + 3 * 1 + (side == 0 ? 0 : 2); // neighbour in -x / +x
  if (d.nbr[chunk * NBSIZE + dir] != NB_NONE)
    return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.M[0] * g.M[1])
    return;
  const int iz = t / g.M[1], iy = t - iz * g.M[1];
  const int Nb = g.nb, Lbx = g.Lb[2], Ubx = g.Ub[2];
  double*   uf = d.uf + (int64_t)chunk * g.Ng * 6;
  const bool estag = iz <= g.M[0] - 2 && iy <= g.M[1] - 2; // range of the normal E component
  const bool bstag = iz >= 1 && iy >= 1;                   // range of the normal B component

  if (side == 0 && kind == PICNIX_BC_WALL) {
    for (int ix = 0; ix < 2 * Nb; ix++) {
      const int ix1 = Lbx - ix + Nb - 1, ix2 = Lbx + Nb;
      F(uf, g, iz, iy, ix1, 1) = 0;
      F(uf, g, iz, iy, ix1, 2) = 0;
      if (estag)
        F(uf, g, iz, iy, ix1, 0) = 0;
      F(uf, g, iz, iy, ix1, 4) = F(uf, g, iz, iy, ix2, 4);
      F(uf, g, iz, iy, ix1, 5) = F(uf, g, iz, iy, ix2, 5);
      if (bstag)
        F(uf, g, iz, iy, ix1, 3) = F(uf, g, iz, iy, ix2, 3);
    }
  } else if (side == 1 && kind == PICNIX_BC_INFLOW) {
    const double* v = g.bc_val[2][1];
    for (int ix = 0; ix < 2 * Nb; ix++) {
      const int ix1 = Ubx + ix - Nb + 1;
      F(uf, g, iz, iy, ix1, 1) = v[1];
      F(uf, g, iz, iy, ix1, 2) = v[2];
      if (bstag)
        F(uf, g, iz, iy, ix1, 3) = v[3];
    }
    for (int ix = 0; ix < 2 * Nb - 1; ix++) {
      const int ix1 = Ubx + ix - Nb + 2;
      if (estag)
        F(uf, g, iz, iy, ix1, 0) = v[0];
      F(uf, g, iz, iy, ix1, 4) = v[4];
      F(uf, g, iz, iy, ix1, 5) = v[5];
    }
  }
}

} // namespace

// set_boundary_field(mode) for all local chunks; called at the end of launch_halo_end(mode)
int launch_boundary_field(picnix_arena* a, int mode)
{
  if (!a->any_bc || mode != PICNIX_BOUNDARY_EMF)
    return PICNIX_OK;
  const Geom& g = a->g;
  for (int side = 0; side < 2; side++) {
    if (g.bc_kind[1][side] == PICNIX_BC_CONDUCTING) {
      dim3 grid((g.M[0] * g.M[2] + BC_THREADS - 1) / BC_THREADS, g.nchunk);
      conducting_y_kernel<<<grid, BC_THREADS, 0, a->stream>>>(g, a->d, side, 0);
      conducting_y_kernel<<<grid, BC_THREADS, 0, a->stream>>>(g, a->d, side, 1);
      a->kernel_launches += 2;
    }
    if (g.bc_kind[2][side] == PICNIX_BC_WALL || g.bc_kind[2][side] == PICNIX_BC_INFLOW) {
      dim3 grid((g.M[0] * g.M[1] + BC_THREADS - 1) / BC_THREADS, g.nchunk);
      shock_x_kernel<<<grid, BC_THREADS, 0, a->stream>>>(g, a->d, side, g.bc_kind[2][side]);
      a->kernel_launches++;
    }
  }
  return check_cuda(a, cudaGetLastError(), "boundary_field");
}

} // namespace picnix

using namespace picnix;

extern "C" int picnix_cuda_set_boundary_condition(picnix_arena_t* a, int32_t axis, int32_t side, int32_t kind,
                                                   const double* values)
{
  if (a == nullptr || axis < 0 || axis > 2 || side < 0 || side > 1)
    return PICNIX_ERR_INVALID;
  if (kind != PICNIX_BC_NONE && a->cfg.periodic[axis])
    return fail(a, PICNIX_ERR_INVALID, "a physical boundary condition needs a non-periodic direction");
  const bool ok = kind == PICNIX_BC_NONE || (kind == PICNIX_BC_CONDUCTING && axis == 1) ||
                  (kind == PICNIX_BC_WALL && axis == 2 && side == 0) ||
                  (kind == PICNIX_BC_INFLOW && axis == 2 && side == 1);
  if (!ok)
    return fail(a, PICNIX_ERR_INVALID,
                "boundary kind not available on this face: CONDUCTING is implemented for walls normal to y, "
                "WALL for the lower and INFLOW for the upper x boundary (the reference's mrx and shock problems)");
  a->g.bc_kind[axis][side] = kind;
  for (int k = 0; k < 6; k++)
    a->g.bc_val[axis][side][k] = (values != nullptr && kind == PICNIX_BC_INFLOW) ? values[k] : 0.0;
  a->any_bc = false;
  for (int i = 0; i < 3; i++)
    for (int s = 0; s < 2; s++)
      a->any_bc = a->any_bc || a->g.bc_kind[i][s] != PICNIX_BC_NONE;
  a->g.any_particle_bc = 0;
  for (int i = 0; i < 3; i++)
    for (int s = 0; s < 2; s++)
      if (a->g.bc_kind[i][s] == PICNIX_BC_CONDUCTING || a->g.bc_kind[i][s] == PICNIX_BC_WALL)
        a->g.any_particle_bc = 1;
  return PICNIX_OK;
}
