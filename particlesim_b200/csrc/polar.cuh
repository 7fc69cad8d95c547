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
#include "traverse.cuh"

namespace psim {

constexpr uint32_t kPolarHasDipole = 1u << 8;

// records in cell order: A = {x, y, charge, radius with the sign bit set if the body has a dipole}, B = {species | flags,
// body index, e.rel_pos}
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
    recA[k] = make_float4(p.x, p.y, p.z, dip ? __uint_as_float(__float_as_uint(p.w) | 0x80000000u) : p.w);
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

// Both directions of a dipole-dipole pair at once.  The four site-to-site separations (nucleus / electron of A against
// nucleus / electron of B) are the same whichever body plays "i": d changes sign, dist, min_sep (a + b = b + a) and the
// denominator do not.  So the square roots and denominators (IEEE) or reciprocals (fast) are taken once per site pair
// and the two directions only differ in the source charge - term by term the values polar_pair_force gives.
struct SitePair {
  float dx, dy, w;  // d = point - source; w = denominator (IEEE) or its reciprocal (fast); w = 0: no field (r_eff == 0)
};
template <bool IEEE>
__device__ __forceinline__ SitePair site_pair(float px, float py, float pr, float sx, float sy, float sr, float eps_sq) {
  SitePair g;
  if (!IEEE) {
    g.dx = px - sx, g.dy = py - sy;
    const float d2 = fmaf(g.dx, g.dx, g.dy * g.dy);
    const float dist = d2 > 0.0f ? d2 * rsqrtf(d2) : 0.0f;
    const float r_eff = fmaxf(dist, pr + sr);
    g.w = r_eff == 0.0f ? 0.0f : __frcp_rn(fmaf(r_eff, r_eff, eps_sq) * r_eff);
    return g;
  }
  g.dx = __fsub_rn(px, sx), g.dy = __fsub_rn(py, sy);
  const float dist = __fsqrt_rn(__fadd_rn(__fmul_rn(g.dx, g.dx), __fmul_rn(g.dy, g.dy)));
  const float r_eff = fmaxf(dist, __fadd_rn(pr, sr));
  g.w = r_eff == 0.0f ? 0.0f : __fmul_rn(__fadd_rn(__fmul_rn(r_eff, r_eff), eps_sq), r_eff);
  return g;
}
// field at the pair's point from charge q at its source (sign = +1), or at its source from q at its point (-1)
template <bool IEEE>
__device__ __forceinline__ float2 site_field(const SitePair& g, float sign, float q, float k_e) {
  if (fabsf(q) < 1.1920929e-07f || g.w == 0.0f) return make_float2(0.f, 0.f);
  if (!IEEE) {
    const float s = sign * (k_e * q) * g.w;
    return make_float2(g.dx * s, g.dy * s);
  }
  const float s = __fdiv_rn(__fmul_rn(k_e, q), g.w);
  return make_float2(sign * __fmul_rn(g.dx, s), sign * __fmul_rn(g.dy, s));
}
// A and B both EC / DMC with an electron, ConjugatePair model: f0 = force on A as the polar body (B its neighbour),
// f1 = force on B as the polar body (A its neighbour); ok0 / ok1 false where the reference skips (all fields zero)
template <bool IEEE>
__device__ __forceinline__ void polar_pair_both(float ax_, float ay_, float ar, float aq, float arx, float ary, float a_qeff,
                                                float bx, float by, float br, float bq, float brx, float bry, float b_qeff,
                                                const PolarParams& P, float2& f0, bool& ok0, float2& f1, bool& ok1) {
  const float aex = __fadd_rn(ax_, arx), aey = __fadd_rn(ay_, ary), bex = __fadd_rn(bx, brx), bey = __fadd_rn(by, bry);
  const SitePair NN = site_pair<IEEE>(ax_, ay_, ar, bx, by, br, P.epsilon_sq);      // A nucleus  <- B nucleus
  const SitePair EN = site_pair<IEEE>(aex, aey, 0.0f, bx, by, br, P.epsilon_sq);    // A electron <- B nucleus
  const SitePair NE = site_pair<IEEE>(ax_, ay_, ar, bex, bey, 0.0f, P.epsilon_sq);  // A nucleus  <- B electron
  const SitePair EE = site_pair<IEEE>(aex, aey, 0.0f, bex, bey, 0.0f, P.epsilon_sq);
  auto add = [](float2& a, const float2 t) { a.x = __fadd_rn(a.x, t.x), a.y = __fadd_rn(a.y, t.y); };
  auto subt = [](float2& a, const float2 t) { a.x = __fsub_rn(a.x, t.x), a.y = __fsub_rn(a.y, t.y); };
  {  // A as i (forces.rs:99-158 with i = A, j = B)
    float2 fn = make_float2(0.f, 0.f), fe = make_float2(0.f, 0.f);
    add(fn, site_field<IEEE>(NN, 1.0f, bq, P.k_e));
    add(fe, site_field<IEEE>(EN, 1.0f, bq, P.k_e));
    add(fn, site_field<IEEE>(NN, 1.0f, b_qeff, P.k_e));
    subt(fn, site_field<IEEE>(NE, 1.0f, b_qeff, P.k_e));
    add(fe, site_field<IEEE>(EN, 1.0f, b_qeff, P.k_e));
    subt(fe, site_field<IEEE>(EE, 1.0f, b_qeff, P.k_e));
    ok0 = !(fn.x == 0.0f && fn.y == 0.0f && fe.x == 0.0f && fe.y == 0.0f);
    f0 = make_float2(__fmul_rn(__fsub_rn(fn.x, fe.x), a_qeff), __fmul_rn(__fsub_rn(fn.y, fe.y), a_qeff));
  }
  {  // B as i, A as j: the same site pairs seen from the other end
    float2 fn = make_float2(0.f, 0.f), fe = make_float2(0.f, 0.f);
    add(fn, site_field<IEEE>(NN, -1.0f, aq, P.k_e));       // B nucleus  <- A nucleus
    add(fe, site_field<IEEE>(NE, -1.0f, aq, P.k_e));       // B electron <- A nucleus
    add(fn, site_field<IEEE>(NN, -1.0f, a_qeff, P.k_e));
    subt(fn, site_field<IEEE>(EN, -1.0f, a_qeff, P.k_e));  // B nucleus  <- A electron
    add(fe, site_field<IEEE>(NE, -1.0f, a_qeff, P.k_e));
    subt(fe, site_field<IEEE>(EE, -1.0f, a_qeff, P.k_e));  // B electron <- A electron
    ok1 = !(fn.x == 0.0f && fn.y == 0.0f && fe.x == 0.0f && fe.y == 0.0f);
    f1 = make_float2(__fmul_rn(__fsub_rn(fn.x, fe.x), b_qeff), __fmul_rn(__fsub_rn(fn.y, fe.y), b_qeff));
  }
}

// Two passes per body so that the lanes of a warp stay together: (1) a cheap sweep over the candidates of the
// 3 x 3 (or wider) cell block that only measures distances and notes the few that are inside either partner's
// cutoff (~1 in 7 at the reference's densities), (2) the pair forces of the noted candidates - six softened
// point-source fields with IEEE sqrt and divisions each - in cell order, as a dense loop of similar length in every
// lane.  The candidate list lives in shared memory (kPolarList slots per thread) and is drained whenever it is full,
// so the order of the additions is the cell order whatever its capacity.
constexpr int kPolarThreads = 128;
constexpr int kPolarList = 24;

template <bool IEEE>
__global__ void __launch_bounds__(kPolarThreads)
    polar_forces_kernel(const float4* __restrict__ pqr, const uint8_t* __restrict__ species,
                        const uint8_t* __restrict__ ecount, const uint32_t* __restrict__ eoff,
                        const float2* __restrict__ erel, const SpeciesRow* __restrict__ table_g,
                        uint32_t first, uint32_t n, const uint32_t* __restrict__ cell_off,
                        const float4* __restrict__ recA,
                        const float4* __restrict__ recB, const uint32_t* __restrict__ body_cell,
                        const uint32_t* __restrict__ max_cutoff_bits, PolarParams P,
                        float4* __restrict__ acc_mass) {
  __shared__ float s_polar_charge[kMaxSpecies];
  __shared__ uint32_t s_list[kPolarList][kPolarThreads];
  if (threadIdx.x < kMaxSpecies) s_polar_charge[threadIdx.x] = table_g[threadIdx.x].polar_charge;
  __syncthreads();
  const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
  const float max_cutoff = __uint_as_float(*max_cutoff_bits);
  if (!(max_cutoff > 0.0f)) return;  // no polar body with an electron: the reference loop does nothing
  const bool live = i < n;
  const float4 me = live ? pqr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const uint32_t msp = live && species[i] < kMaxSpecies ? species[i] : 0;
  const bool me_dip = live && (msp == 4u || msp == 5u) && ecount[i] != 0;
  float2 mrel = make_float2(0.f, 0.f);
  if (me_dip) mrel = erel[eoff[i]];
  const float me_qeff = s_polar_charge[msp];
  const float my_cut = __fmul_rn(3.0f, me.w), my_cut_sq = __fmul_rn(my_cut, my_cut);
  float4 am = live ? acc_mass[i] : make_float4(0.f, 0.f, 0.f, 1.f);
  const float inv_mass = __frcp_rn(am.w);
  float ax = 0.0f, ay = 0.0f;
  int filled = 0;

  // fast arithmetic: per-thread source weights of `me` (k_e q with the reference's |q| < EPSILON guard folded in; the
  // dipole terms only exist in the ConjugatePair model)
  const float kEps = 1.1920929e-07f;
  const float me_kq = fabsf(me.z) < kEps ? 0.0f : P.k_e * me.z;
  const float me_ke = (me_dip && P.dipole_model == 1 && !(fabsf(me_qeff) < kEps)) ? P.k_e * me_qeff : 0.0f;
  const float mex = __fadd_rn(me.x, mrel.x), mey = __fadd_rn(me.y, mrel.y);

  auto drain = [&]() {
    if constexpr (!IEEE) {
      // One branch-free body for every kind of pair (dipole / dipole, dipole / ion, inside one cutoff or both): the
      // four site-to-site geometries are always taken, a partner without a dipole has its electron site on its nucleus
      // and weight zero there, and the two directions are added under a select.  The lanes of a warp stay together
      // whatever mix of species their lists hold.
      for (int t = 0; t < filled; ++t) {
        const uint32_t k = s_list[t][threadIdx.x];
        const float4 b4 = __ldg(&recB[k]);
        float4 a4 = __ldg(&recA[k]);
        a4.w = fabsf(a4.w);  // the sign bit is the dipole flag of the distance sweep
        const uint32_t jbits = __float_as_uint(b4.x);
        const bool j_dip = (jbits & kPolarHasDipole) != 0;
        uint32_t jsp = jbits & 0xffu;
        if (jsp >= kMaxSpecies) jsp = 0;
        const float j_qeff = s_polar_charge[jsp];
        const float rx = __fsub_rn(a4.x, me.x), ry = __fsub_rn(a4.y, me.y);
        const float r2 = __fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry));
        const float jc = __fmul_rn(3.0f, a4.w);
        const bool other = __float_as_uint(b4.y) != i;           // the body itself is in its own cell
        const bool use0 = other && me_dip && r2 < my_cut_sq;     // me as the polar body i
        const bool use1 = other && j_dip && r2 < __fmul_rn(jc, jc);  // the candidate as i, me its neighbour: reaction
        const float j_kq = fabsf(a4.z) < kEps ? 0.0f : P.k_e * a4.z;
        const float j_ke = (j_dip && P.dipole_model == 1 && !(fabsf(j_qeff) < kEps)) ? P.k_e * j_qeff : 0.0f;
        const float jex = __fadd_rn(a4.x, b4.z), jey = __fadd_rn(a4.y, b4.w);
        // geometry of (point of me) <- (source of j): d = point - source, w = 1 / ((r_eff^2 + eps^2) r_eff)
        auto geom = [&](float px, float py, float sx, float sy, float min_sep, float& dx, float& dy) {
          dx = px - sx, dy = py - sy;
          const float d2 = fmaf(dx, dx, dy * dy);
          const float r_eff = fmaxf(d2 * rsqrt_ftz(d2), min_sep);  // fmaxf drops the NaN of d2 == 0
          const float den = fmaf(r_eff, r_eff, P.epsilon_sq) * r_eff;
          return den >= 1.17549435e-38f ? rcp_ftz(den) : 0.0f;     // coincident zero-radius sites: no field
        };
        float nnx, nny, enx, eny, nex, ney, eex, eey;
        const float wNN = geom(me.x, me.y, a4.x, a4.y, me.w + a4.w, nnx, nny);  // my nucleus  <- its nucleus
        const float wEN = geom(mex, mey, a4.x, a4.y, a4.w, enx, eny);           // my electron <- its nucleus
        const float wNE = geom(me.x, me.y, jex, jey, me.w, nex, ney);           // my nucleus  <- its electron
        const float wEE = geom(mex, mey, jex, jey, 0.0f, eex, eey);             // my electron <- its electron
        // me as i (forces.rs:99-158): (field at my nucleus - field at my electron) * my q_eff
        const float cNN0 = (j_kq + j_ke) * wNN, cEN0 = (j_kq + j_ke) * wEN, cNE0 = j_ke * wNE, cEE0 = j_ke * wEE;
        const float f0x = (fmaf(cNN0, nnx, -cNE0 * nex) - fmaf(cEN0, enx, -cEE0 * eex)) * me_qeff;
        const float f0y = (fmaf(cNN0, nny, -cNE0 * ney) - fmaf(cEN0, eny, -cEE0 * eey)) * me_qeff;
        // the candidate as i: the same site pairs seen from the other end (d changes sign)
        const float cNN1 = (me_kq + me_ke) * wNN, cNE1 = (me_kq + me_ke) * wNE, cEN1 = me_ke * wEN, cEE1 = me_ke * wEE;
        const float f1x = (fmaf(cNE1, nex, -cEE1 * eex) - fmaf(cNN1, nnx, -cEN1 * enx)) * j_qeff;
        const float f1y = (fmaf(cNE1, ney, -cEE1 * eey) - fmaf(cNN1, nny, -cEN1 * eny)) * j_qeff;
        float tx = use0 ? f0x : 0.0f, ty = use0 ? f0y : 0.0f;
        tx -= use1 ? f1x : 0.0f, ty -= use1 ? f1y : 0.0f;
        ax = fmaf(tx, inv_mass, ax);
        ay = fmaf(ty, inv_mass, ay);
      }
      filled = 0;
      return;
    } else {
    for (int t = 0; t < filled; ++t) {
      const uint32_t k = s_list[t][threadIdx.x];
      const float4 b4 = __ldg(&recB[k]);
      float4 a4 = __ldg(&recA[k]);
      a4.w = fabsf(a4.w);  // the sign bit is the dipole flag of the distance sweep
      const uint32_t jbits = __float_as_uint(b4.x);
      const bool j_dip = (jbits & kPolarHasDipole) != 0;
      if (__float_as_uint(b4.y) == i || (!me_dip && !j_dip)) continue;  // the body itself; ion / ion: no polar term
      const float rx = __fsub_rn(a4.x, me.x), ry = __fsub_rn(a4.y, me.y);
      const float r2 = __fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry));
      uint32_t jsp = jbits & 0xffu;
      if (jsp >= kMaxSpecies) jsp = 0;
      const float j_qeff_tab = s_polar_charge[jsp];
      float fx, fy;
      const float jc = __fmul_rn(3.0f, a4.w);
      if (me_dip && j_dip && P.dipole_model == 1 && r2 < my_cut_sq && r2 < __fmul_rn(jc, jc)) {
        // dipole against dipole inside both cutoffs (the bulk of the work): both directions from shared geometry
        float2 f0, f1;
        bool ok0, ok1;
        polar_pair_both<IEEE>(me.x, me.y, me.w, me.z, mrel.x, mrel.y, me_qeff, a4.x, a4.y, a4.w, a4.z, b4.z, b4.w,
                              j_qeff_tab, P, f0, ok0, f1, ok1);
        if (ok0) {
          ax = __fadd_rn(ax, IEEE ? __fdiv_rn(f0.x, am.w) : f0.x * inv_mass);
          ay = __fadd_rn(ay, IEEE ? __fdiv_rn(f0.y, am.w) : f0.y * inv_mass);
        }
        if (ok1) {
          ax = __fsub_rn(ax, IEEE ? __fdiv_rn(f1.x, am.w) : f1.x * inv_mass);
          ay = __fsub_rn(ay, IEEE ? __fdiv_rn(f1.y, am.w) : f1.y * inv_mass);
        }
        continue;
      }
      // me as the polar body i, the candidate as its neighbour j
      if (me_dip && r2 < my_cut_sq) {
        if (polar_pair_force<IEEE>(me.x, me.y, me.w, mrel.x, mrel.y, me_qeff, a4.x, a4.y, a4.w, a4.z, j_dip, b4.z, b4.w,
                             j_dip ? j_qeff_tab : 0.0f, P, fx, fy)) {
          ax = __fadd_rn(ax, IEEE ? __fdiv_rn(fx, am.w) : fx * inv_mass);
          ay = __fadd_rn(ay, IEEE ? __fdiv_rn(fy, am.w) : fy * inv_mass);
        }
      }
      // the candidate as the polar body i, me as its neighbour j: reaction  -force / m_me
      if (j_dip) {
        if (r2 < __fmul_rn(jc, jc)) {
          if (polar_pair_force<IEEE>(a4.x, a4.y, a4.w, b4.z, b4.w, j_qeff_tab, me.x, me.y, me.w, me.z, me_dip, mrel.x, mrel.y,
                               me_dip ? me_qeff : 0.0f, P, fx, fy)) {
            ax = __fsub_rn(ax, IEEE ? __fdiv_rn(fx, am.w) : fx * inv_mass);
            ay = __fsub_rn(ay, IEEE ? __fdiv_rn(fy, am.w) : fy * inv_mass);
          }
        }
      }
    }
    filled = 0;
    }
  };

  if (live) {
    const uint32_t c = body_cell[i];
    const int cx = (int)(c % P.g.gx), cy = (int)(c / P.g.gx);
    const int range = (int)ceilf(max_cutoff / P.g.cell_size);
    const int y0 = max(cy - range, 0), y1 = min(cy + range, (int)P.g.gy - 1);
    const int x0 = max(cx - range, 0), x1 = min(cx + range, (int)P.g.gx - 1);
    // any pair closer than this may interact (either partner's 3 * radius, bounded by the largest one present)
    const float any_cut_sq = __fmul_rn(max_cutoff, max_cutoff);
    for (int y = y0; y <= y1; ++y) {
      // the cells x0 .. x1 of a row are one contiguous run of the cell order (same candidate order as cell by cell)
      const uint32_t row = (uint32_t)y * P.g.gx;
      const uint32_t k1 = cell_off[row + (uint32_t)x1 + 1u];
      for (uint32_t kb = cell_off[row + (uint32_t)x0]; kb < k1; kb += 4) {
        // four candidates' records in flight at a time (same candidate order)
        float4 a4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) a4[u] = __ldg(&recA[kb + u < k1 ? kb + u : k1 - 1]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t k = kb + u;
          const float rx = __fsub_rn(a4[u].x, me.x), ry = __fsub_rn(a4[u].y, me.y);
          const float r2 = __fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry));
          if (k >= k1 || !(r2 < any_cut_sq)) continue;
          // A body without a dipole only interacts with dipoles (the flag rides on the sign of record A's radius, so
          // this loop never reads record B); a dipole takes every candidate in range, and the body itself is dropped
          // in the second pass.
          if (!me_dip && !(__float_as_uint(a4[u].w) >> 31)) continue;
          s_list[filled][threadIdx.x] = k;
          if (++filled == kPolarList) drain();
        }
      }
    }
  }
  drain();
  if (live) {
    am.x = __fadd_rn(am.x, ax);
    am.y = __fadd_rn(am.y, ay);
    acc_mass[i] = am;
  }
}

}  // namespace psim
