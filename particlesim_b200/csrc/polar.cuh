// polar.cuh — polarization forces between polar solvent molecules and their neighbours.
//
// Replaces forces::apply_polar_forces (src/simulation/forces.rs:52-175), the pass that sits between
// `attract` and the LJ pass in Simulation::step (simulation.rs:1007).  The reference runs it as a
// serial loop over bodies i (EC / DMC with a bound electron), asks the cell list for the neighbours
// within 3 * radius_i and adds  force / m_i  to i and  -force / m_j  to every neighbour j.  Here it is
// a gather: each body sums the forces it exerts as "i" and the reactions it receives as "j" (the
// pair term is evaluated by both partners, so no atomics and a deterministic result).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cells.cuh"

namespace psim {

constexpr uint32_t kPolarHasDipole = 1u << 8;

// records in cell order: A = {x, y, charge, radius}, B = {species | flags, body index, e.rel_pos}
__global__ void __launch_bounds__(256)
    polar_records_kernel(const uint32_t* __restrict__ order, uint32_t n, const float4* __restrict__ pqr,
                         const uint8_t* __restrict__ species, const uint8_t* __restrict__ ecount,
                         const uint32_t* __restrict__ eoff, const float2* __restrict__ erel,
                         float4* __restrict__ recA, float4* __restrict__ recB,
                         uint32_t* __restrict__ max_cutoff_bits) {
  float local_max = 0.0f;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const uint32_t b = order[k];
    const float4 p = pqr[b];
    const uint32_t sp = species[b];
    const bool dip = (sp == 4u || sp == 5u) && ecount[b] != 0;  // EC | DMC with an electron (forces.rs:67-72)
    float2 r = make_float2(0.f, 0.f);
    if (dip) {
      r = erel[eoff[b]];
      local_max = fmaxf(local_max, __fmul_rn(3.0f, p.w));
    }
    recA[k] = p;
    recB[k] = make_float4(__uint_as_float(sp | (dip ? kPolarHasDipole : 0u)), __uint_as_float(b), r.x, r.y);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, off));
  if ((threadIdx.x & 31) == 0 && local_max > 0.0f) atomicMax(max_cutoff_bits, __float_as_uint(local_max));
}

struct PolarParams {
  GridDims g;
  float k_e, epsilon_sq;
  int dipole_model;  // 0 SingleOffset, 1 ConjugatePair (config.rs:270-281)
};

// field_from_source closure of forces.rs:84-97.  IEEE: the reference's operations one by one (IEEE sqrt and division,
// no contraction); otherwise (psim_config.parity_mode 0, like the traversal) MUFU rsqrt / rcp and FMA on the same terms.
template <bool IEEE>
__device__ __forceinline__ float2 polar_field(float px, float py, float point_radius, float sx, float sy,
                                              float src_radius, float src_charge, float k_e, float eps_sq) {
  if (fabsf(src_charge) < 1.1920929e-07f) return make_float2(0.f, 0.f);
  if (!IEEE) {
    const float dx = px - sx, dy = py - sy;
    const float d2 = fmaf(dx, dx, dy * dy);
    const float dist = d2 > 0.0f ? d2 * rsqrtf(d2) : 0.0f;
    const float r_eff = fmaxf(dist, point_radius + src_radius);
    if (r_eff == 0.0f) return make_float2(0.f, 0.f);
    const float s = __fdividef(k_e * src_charge, fmaf(r_eff, r_eff, eps_sq) * r_eff);
    return make_float2(dx * s, dy * s);
  }
  const float dx = __fsub_rn(px, sx), dy = __fsub_rn(py, sy);
  const float dist = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
  const float r_eff = fmaxf(dist, __fadd_rn(point_radius, src_radius));
  // Two zero-radius sites at exactly the same place (an electron of i on an electron of j): the reference evaluates
  // d * (k q / 0) = 0 * inf = NaN there, which turns every node centre of its next build into NaN.  With f32
  // positions on a ~1e-3 A grid (|x| ~ 8000 A) that happens a few times per step at 16 M bodies, so the term is
  // taken as the zero vector d already is (documented divergence).
  if (r_eff == 0.0f) return make_float2(0.f, 0.f);
  const float denom = __fmul_rn(__fadd_rn(__fmul_rn(r_eff, r_eff), eps_sq), r_eff);
  const float s = __fdiv_rn(__fmul_rn(k_e, src_charge), denom);
  return make_float2(__fmul_rn(dx, s), __fmul_rn(dy, s));
}

// force on polar body i (nucleus at ipos with radius irad, electron at ipos + irel) from neighbour j
// (forces.rs:99-158); returns false if the reference skips the pair (both fields exactly zero)
template <bool IEEE>
__device__ __forceinline__ bool polar_pair_force(float ix, float iy, float irad, float irx, float iry, float i_qeff,
                                                 float jx, float jy, float jrad, float jq, bool j_dip,
                                                 float jrx, float jry, float j_qeff, const PolarParams& P,
                                                 float& fx, float& fy) {
  const float iex = __fadd_rn(ix, irx), iey = __fadd_rn(iy, iry);
  float2 fn = polar_field<IEEE>(ix, iy, irad, jx, jy, jrad, jq, P.k_e, P.epsilon_sq);
  float2 fe = polar_field<IEEE>(iex, iey, 0.0f, jx, jy, jrad, jq, P.k_e, P.epsilon_sq);
  // the reference starts from Vec2::zero() and += each term
  fn.x = __fadd_rn(0.0f, fn.x), fn.y = __fadd_rn(0.0f, fn.y);
  fe.x = __fadd_rn(0.0f, fe.x), fe.y = __fadd_rn(0.0f, fe.y);
  if (P.dipole_model == 1 && j_dip) {
    const float jex = __fadd_rn(jx, jrx), jey = __fadd_rn(jy, jry);
    float2 t = polar_field<IEEE>(ix, iy, irad, jx, jy, jrad, j_qeff, P.k_e, P.epsilon_sq);
    fn.x = __fadd_rn(fn.x, t.x), fn.y = __fadd_rn(fn.y, t.y);
    t = polar_field<IEEE>(ix, iy, irad, jex, jey, 0.0f, j_qeff, P.k_e, P.epsilon_sq);
    fn.x = __fsub_rn(fn.x, t.x), fn.y = __fsub_rn(fn.y, t.y);
    t = polar_field<IEEE>(iex, iey, 0.0f, jx, jy, jrad, j_qeff, P.k_e, P.epsilon_sq);
    fe.x = __fadd_rn(fe.x, t.x), fe.y = __fadd_rn(fe.y, t.y);
    t = polar_field<IEEE>(iex, iey, 0.0f, jex, jey, 0.0f, j_qeff, P.k_e, P.epsilon_sq);
    fe.x = __fsub_rn(fe.x, t.x), fe.y = __fsub_rn(fe.y, t.y);
  }
  if (fn.x == 0.0f && fn.y == 0.0f && fe.x == 0.0f && fe.y == 0.0f) return false;
  fx = __fmul_rn(__fsub_rn(fn.x, fe.x), i_qeff);
  fy = __fmul_rn(__fsub_rn(fn.y, fe.y), i_qeff);
  return true;
}

// Three steps per batch, so that the expensive part runs on full warps:
//   (1) every lane sweeps the candidates of its body's 3 x 3 (or wider) cell block at its own pace, only measuring
//       distances, and notes the few that are inside the largest cutoff present in its shared-memory list;
//   (2) the warp pools the lists: each noted (body, candidate) pair is two work items - the body as the polar body i,
//       and the candidate as i with the body taking the reaction - and the lanes take items round-robin, whoever noted
//       them; an item is one pair force (two to six softened point-source fields) written to the owner's result slot;
//   (3) every lane adds up its own slots in list order, so a body's sum has the cell order of the reference's
//       neighbour query whichever lane evaluated the terms: no atomics, deterministic, same bits as a private loop.
// A list holds kPolarList candidates; a lane whose list is full waits for the warp's next drain.
constexpr int kPolarThreads = 128;
constexpr int kPolarList = 16;

struct PolarMe {  // a lane's own body, readable by the whole warp
  float x, y, q, r, relx, rely, qeff, inv_mass, mass;
  uint32_t dip;
};

template <bool IEEE>
__global__ void __launch_bounds__(kPolarThreads)
    polar_forces_kernel(const float4* __restrict__ pqr, const uint8_t* __restrict__ species,
                        const uint8_t* __restrict__ ecount, const uint32_t* __restrict__ eoff,
                        const float2* __restrict__ erel, const SpeciesRow* __restrict__ table_g,
                        uint32_t first, uint32_t n, const uint32_t* __restrict__ cell_start,
                        const uint32_t* __restrict__ cell_end, const float4* __restrict__ recA,
                        const float4* __restrict__ recB, const uint32_t* __restrict__ body_cell,
                        const uint32_t* __restrict__ max_cutoff_bits, PolarParams P,
                        float4* __restrict__ acc_mass) {
  __shared__ float s_polar_charge[kMaxSpecies];
  __shared__ uint32_t s_list[kPolarList][kPolarThreads];
  __shared__ float2 s_res[2 * kPolarList][kPolarThreads];
  __shared__ PolarMe s_me[kPolarThreads];
  __shared__ int s_start[kPolarThreads];
  if (threadIdx.x < kMaxSpecies) s_polar_charge[threadIdx.x] = table_g[threadIdx.x].polar_charge;
  __syncthreads();
  const uint32_t FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
  const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
  const float max_cutoff = __uint_as_float(*max_cutoff_bits);
  if (!(max_cutoff > 0.0f)) return;  // no polar body with an electron: the reference loop does nothing
  const bool live = i < n;
  const float4 me = live ? pqr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const uint32_t msp = live && species[i] < kMaxSpecies ? species[i] : 0;
  const bool me_dip = live && (msp == 4u || msp == 5u) && ecount[i] != 0;
  float2 mrel = make_float2(0.f, 0.f);
  if (me_dip) mrel = erel[eoff[i]];
  float4 am = live ? acc_mass[i] : make_float4(0.f, 0.f, 0.f, 1.f);
  {
    PolarMe m;
    m.x = me.x, m.y = me.y, m.q = me.z, m.r = me.w, m.relx = mrel.x, m.rely = mrel.y, m.qeff = s_polar_charge[msp];
    m.inv_mass = __frcp_rn(am.w), m.mass = am.w, m.dip = me_dip ? 1u : 0u;
    s_me[threadIdx.x] = m;
  }
  __syncwarp();
  float ax = 0.0f, ay = 0.0f;
  int filled = 0;

  // steps (2) and (3) for what the lanes of this warp have noted so far; called by all 32 lanes together
  auto drain = [&]() {
    int incl = 2 * filled;  // items of lane l occupy [start_l, start_l + 2 * filled_l)
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(FULL, incl, off);
      if (lane >= off) incl += v;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    s_start[threadIdx.x] = incl - 2 * filled;
    __syncwarp();
    for (int t = lane; t < total; t += 32) {
      int owner = 0;  // the last lane whose start is <= t (lanes without items share their successor's start)
#pragma unroll
      for (int step = 16; step > 0; step >>= 1)
        if (s_start[wbase + owner + step] <= t) owner += step;
      const int item = t - s_start[wbase + owner], slot = item >> 1, dir = item & 1;
      const uint32_t k = s_list[slot][wbase + owner];
      const PolarMe m = s_me[wbase + owner];
      const float4 a4 = __ldg(&recA[k]);
      const float4 b4 = __ldg(&recB[k]);
      const uint32_t jbits = __float_as_uint(b4.x);
      const bool j_dip = (jbits & kPolarHasDipole) != 0;
      uint32_t jsp = jbits & 0xffu;
      if (jsp >= kMaxSpecies) jsp = 0;
      const float j_qeff = s_polar_charge[jsp];
      const float rx = __fsub_rn(a4.x, m.x), ry = __fsub_rn(a4.y, m.y);
      const float r2 = __fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry));
      float fx = 0.0f, fy = 0.0f;
      float2 out = make_float2(0.0f, 0.0f);
      if (dir == 0) {
        // the owner as the polar body i, the candidate as its neighbour j
        const float cut = __fmul_rn(3.0f, m.r);
        if (m.dip && r2 < __fmul_rn(cut, cut) &&
            polar_pair_force<IEEE>(m.x, m.y, m.r, m.relx, m.rely, m.qeff, a4.x, a4.y, a4.w, a4.z, j_dip, b4.z, b4.w,
                                   j_dip ? j_qeff : 0.0f, P, fx, fy))
          out = make_float2(IEEE ? __fdiv_rn(fx, m.mass) : fx * m.inv_mass, IEEE ? __fdiv_rn(fy, m.mass) : fy * m.inv_mass);
      } else {
        // the candidate as the polar body i, the owner as its neighbour j: reaction  -force / m_owner
        const float jc = __fmul_rn(3.0f, a4.w);
        if (j_dip && r2 < __fmul_rn(jc, jc) &&
            polar_pair_force<IEEE>(a4.x, a4.y, a4.w, b4.z, b4.w, j_qeff, m.x, m.y, m.r, m.q, m.dip != 0, m.relx, m.rely,
                                   m.dip ? m.qeff : 0.0f, P, fx, fy))
          out = make_float2(-(IEEE ? __fdiv_rn(fx, m.mass) : fx * m.inv_mass), -(IEEE ? __fdiv_rn(fy, m.mass) : fy * m.inv_mass));
      }
      s_res[item][wbase + owner] = out;
    }
    __syncwarp();
    for (int t = 0; t < 2 * filled; ++t) {
      const float2 r = s_res[t][threadIdx.x];
      ax = __fadd_rn(ax, r.x), ay = __fadd_rn(ay, r.y);
    }
    filled = 0;
    __syncwarp();
  };

  // step (1): each lane scans on its own until its list is full or its cells are exhausted, then the warp drains
  uint32_t k = 0, k1 = 0;
  int x0 = 0, x1 = -1, y1 = -1, xx = 0, yy = 0;
  float any_cut_sq = 0.0f;
  bool scanning = live;
  if (live) {
    const uint32_t c = body_cell[i];
    const int cx = (int)(c % P.g.gx), cy = (int)(c / P.g.gx);
    const int range = (int)ceilf(max_cutoff / P.g.cell_size);
    const int y0 = max(cy - range, 0);
    y1 = min(cy + range, (int)P.g.gy - 1);
    x0 = max(cx - range, 0), x1 = min(cx + range, (int)P.g.gx - 1);
    // any pair closer than this may interact (either partner's 3 * radius, bounded by the largest one present)
    any_cut_sq = __fmul_rn(max_cutoff, max_cutoff);
    xx = x0, yy = y0;
    const uint32_t cc = (uint32_t)xx + (uint32_t)yy * P.g.gx;
    k = cell_start[cc], k1 = cell_end[cc];
  }
  do {
    while (scanning && filled < kPolarList) {
      if (k < k1) {
        const float4 a4 = __ldg(&recA[k]);
        const float rx = __fsub_rn(a4.x, me.x), ry = __fsub_rn(a4.y, me.y);
        const float r2 = __fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry));
        if (r2 < any_cut_sq) {
          const float4 b4 = __ldg(&recB[k]);
          if (__float_as_uint(b4.y) != i && (me_dip || (__float_as_uint(b4.x) & kPolarHasDipole))) {
            s_list[filled][threadIdx.x] = k;
            ++filled;
          }
        }
        ++k;
      } else if (xx < x1) {
        ++xx;
        const uint32_t cc = (uint32_t)xx + (uint32_t)yy * P.g.gx;
        k = cell_start[cc], k1 = cell_end[cc];
      } else if (yy < y1) {
        ++yy, xx = x0;
        const uint32_t cc = (uint32_t)xx + (uint32_t)yy * P.g.gx;
        k = cell_start[cc], k1 = cell_end[cc];
      } else {
        scanning = false;
      }
    }
    __syncwarp();
    drain();
  } while (__any_sync(FULL, scanning));
  if (live) {
    am.x = __fadd_rn(am.x, ax);
    am.y = __fadd_rn(am.y, ay);
    acc_mass[i] = am;
  }
}

}  // namespace psim
