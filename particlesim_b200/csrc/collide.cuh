// collide.cuh — one pass of collision::collide (src/simulation/collision.rs:62-372).
//
// The reference finds every pair of overlapping bounding squares with a broccoli BVH and calls resolve(i, j) on
// each, in parallel, on the shared mutable `Simulation` (collision.rs:148-156): the order in which the pairs are
// resolved - and hence what positions a later resolve sees - is whatever the thread pool happens to do.  There is no
// order to be faithful to, so the device takes the one order-free reading of the pass: every pair is resolved from
// the state AT THE START of the pass (resolve's arithmetic, operation for operation: positional separation, the
// time-of-impact rewind with the 1.5 d.v / d^2 impulse, the degenerate-pair and stationary fall-backs, the metal
// stiffness / soft-ion weight modifiers) and each body applies the sum of the changes its pairs ask for.  For a body
// in one overlapping pair this IS resolve(i, j); for a body in several it is the Jacobi form of the reference's
// Gauss-Seidel sweep.  Broad phase: the cell list at a cell size of the largest diameter present, 3 x 3 cells.
// Gather form (each pair is evaluated by both partners), no atomics, deterministic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cells.cuh"

namespace psim {

struct CollideParams {
  GridDims g;
  float correction_scale;   // 1 / num_passes (collision.rs:298)
  float softness;           // li_collision_softness clamped to [0, 1] (config.rs:235)
  uint32_t soft_li, soft_an;  // soft_collision_lithium_ion / soft_collision_anion (config.rs:484-485)
  float domain_depth;
};

// records in cell order: A = {x, y, z, radius}, B = {vx, vy, vz, mass}; species and body index come from cpos
__global__ void __launch_bounds__(256)
    collide_records_kernel(const uint32_t* __restrict__ order, uint32_t n, const float4* __restrict__ pqr,
                           const float4* __restrict__ velz, const float4* __restrict__ accm,
                           float4* __restrict__ recA, float4* __restrict__ recB) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const uint32_t b = order[k];
    const float4 p = pqr[b], v = velz[b];
    recA[k] = make_float4(p.x, p.y, v.z, p.w);
    recB[k] = make_float4(v.x, v.y, v.w, accm[b].w);
  }
}

struct CollideBody {
  float x, y, z, r, vx, vy, vz, m;
  uint32_t species, index;
};
struct CollideDelta {  // what one resolve() does to the FIRST body of the pair
  float dx, dy, dz, dvx, dvy, dvz;
  bool set_pos;        // degenerate fall-back: absolute position / z instead of a displacement
  float px, py, pz;
};

__device__ __forceinline__ bool finitef(float v) { return fabsf(v) < __int_as_float(0x7f800000); }
__device__ __forceinline__ bool collide_is_metal(uint32_t s) { return s == 1u || s == 2u; }

// apply_collision_modifiers (collision.rs:16-60)
__device__ __forceinline__ void collide_weights(uint32_t si, uint32_t sj, float wi, float wj, const CollideParams& P,
                                                float& mi, float& mj) {
  const bool im = collide_is_metal(si), jm = collide_is_metal(sj);
  const float s = P.softness;
  if (im && !jm) {
    mj = __fadd_rn(wj, __fmul_rn(wi, s));
    mi = __fmul_rn(wi, __fsub_rn(1.0f, s));
    return;
  }
  if (jm && !im) {
    mi = __fadd_rn(wi, __fmul_rn(wj, s));
    mj = __fmul_rn(wj, __fsub_rn(1.0f, s));
    return;
  }
  const bool soft = (P.soft_li && (si == 0u || sj == 0u)) || (P.soft_an && (si == 3u || sj == 3u));
  if (soft) {
    const float scale = __fsub_rn(1.0f, s);
    mi = __fmul_rn(wi, scale), mj = __fmul_rn(wj, scale);
    return;
  }
  mi = wi, mj = wj;
}

// resolve(sim, i, j, num_passes) of collision.rs:158-372 for the ordered pair (a = bodies[i], b = bodies[j]), i < j,
// from a snapshot; returns false when the pair does not touch.  `first` selects whose change is returned.
__device__ __forceinline__ bool collide_resolve(const CollideBody& a, const CollideBody& b, bool first,
                                                const CollideParams& P, CollideDelta& out) {
  out.dx = out.dy = out.dz = out.dvx = out.dvy = out.dvz = 0.0f;
  out.set_pos = false;
  out.px = out.py = out.pz = 0.0f;
  float dxy_x = __fsub_rn(b.x, a.x), dxy_y = __fsub_rn(b.y, a.y), dz = __fsub_rn(b.z, a.z);
  const float r = __fadd_rn(a.r, b.r);
  float dist_sq = __fadd_rn(__fadd_rn(__fmul_rn(dxy_x, dxy_x), __fmul_rn(dxy_y, dxy_y)), __fmul_rn(dz, dz));
  // non-finite inputs are sanitised to zero by the reference before anything else (collision.rs:175-232); the
  // snapshot records are sanitised by the caller, so only the derived quantity is checked here
  if (!(dist_sq <= __fmul_rn(r, r))) return false;
  const float vx = __fsub_rn(b.vx, a.vx), vy = __fsub_rn(b.vy, a.vy), vz = __fsub_rn(b.vz, a.vz);
  const float d_dot_v = __fadd_rn(__fadd_rn(__fmul_rn(dxy_x, vx), __fmul_rn(dxy_y, vy)), __fmul_rn(dz, vz));
  const float msum = __fadd_rn(a.m, b.m);
  const float w1 = __fdiv_rn(b.m, msum), w2 = __fdiv_rn(a.m, msum);
  float mw1, mw2;
  collide_weights(a.species, b.species, w1, w2, P, mw1, mw2);
  if (d_dot_v >= 0.0f && dist_sq > 0.0f && finitef(dist_sq)) {  // separating or resting: positional correction only
    const float dist = __fsqrt_rn(dist_sq);
    const float corr = __fsub_rn(__fdiv_rn(r, dist), 1.0f);
    const float sx = __fmul_rn(dxy_x, corr), sy = __fmul_rn(dxy_y, corr), sz = __fmul_rn(dz, corr);
    if (first) out.dx = -__fmul_rn(mw1, sx), out.dy = -__fmul_rn(mw1, sy), out.dz = -__fmul_rn(mw1, sz);
    else out.dx = __fmul_rn(mw2, sx), out.dy = __fmul_rn(mw2, sy), out.dz = __fmul_rn(mw2, sz);
    return true;
  }
  const float v_sq = __fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz));
  const float d_sq = dist_sq;
  if (!finitef(d_sq) || d_sq <= 1.0e-8f || !finitef(v_sq)) {
    // deterministic separation along a direction made from the indices (collision.rs:270-296)
    const unsigned long long ji = (unsigned long long)b.index;
    const unsigned long long rot = (ji << 13) | (ji >> 51);
    const float angle = __fmul_rn((float)((unsigned long long)a.index ^ rot), 6.28318530717958647692f / 1024.0f);
    float s, c;
    sincosf(angle, &s, &c);
    const float sep = __fmul_rn(r, 1.001f);
    const float midx = __fmul_rn(__fadd_rn(a.x, b.x), 0.5f), midy = __fmul_rn(__fadd_rn(a.y, b.y), 0.5f);
    float midz = __fmul_rn(__fadd_rn(a.z, b.z), 0.5f);
    midz = fminf(fmaxf(midz, -P.domain_depth), P.domain_depth);
    out.set_pos = true;
    if (first) out.px = __fsub_rn(midx, __fmul_rn(c, __fmul_rn(sep, w1))), out.py = __fsub_rn(midy, __fmul_rn(s, __fmul_rn(sep, w1)));
    else out.px = __fadd_rn(midx, __fmul_rn(c, __fmul_rn(sep, w2))), out.py = __fadd_rn(midy, __fmul_rn(s, __fmul_rn(sep, w2)));
    out.pz = midz;
    return true;
  }
  const float r_sq = __fmul_rn(r, r);
  const float disc = fmaxf(__fsub_rn(__fmul_rn(d_dot_v, d_dot_v), __fmul_rn(v_sq, __fsub_rn(d_sq, r_sq))), 0.0f);
  const float numerator = __fadd_rn(d_dot_v, __fsqrt_rn(disc));
  const float t = __fdiv_rn(__fmul_rn(P.correction_scale, numerator), v_sq);
  if (!finitef(t)) {  // stationary pair: positional correction with the unmodified weights (collision.rs:306-322)
    const float dist = __fsqrt_rn(d_sq);
    if (finitef(dist) && dist > 0.0f) {
      const float corr = __fsub_rn(__fdiv_rn(r, dist), 1.0f);
      const float sx = __fmul_rn(dxy_x, corr), sy = __fmul_rn(dxy_y, corr), sz = __fmul_rn(dz, corr);
      if (first) out.dx = -__fmul_rn(w1, sx), out.dy = -__fmul_rn(w1, sy), out.dz = -__fmul_rn(w1, sz);
      else out.dx = __fmul_rn(w2, sx), out.dy = __fmul_rn(w2, sy), out.dz = __fmul_rn(w2, sz);
    }
    return true;
  }
  // rewind both bodies to the moment of contact, exchange the impulse, advance again (collision.rs:323-360)
  const float p1x = __fsub_rn(a.x, __fmul_rn(a.vx, t)), p1y = __fsub_rn(a.y, __fmul_rn(a.vy, t)), z1 = __fsub_rn(a.z, __fmul_rn(a.vz, t));
  const float p2x = __fsub_rn(b.x, __fmul_rn(b.vx, t)), p2y = __fsub_rn(b.y, __fmul_rn(b.vy, t)), z2 = __fsub_rn(b.z, __fmul_rn(b.vz, t));
  dxy_x = __fsub_rn(p2x, p1x), dxy_y = __fsub_rn(p2y, p1y), dz = __fsub_rn(z2, z1);
  const float ddv = __fadd_rn(__fadd_rn(__fmul_rn(dxy_x, vx), __fmul_rn(dxy_y, vy)), __fmul_rn(dz, vz));
  const float dsq2 = __fadd_rn(__fadd_rn(__fmul_rn(dxy_x, dxy_x), __fmul_rn(dxy_y, dxy_y)), __fmul_rn(dz, dz));
  float scale = (finitef(dsq2) && dsq2 > 0.0f) ? __fdiv_rn(__fmul_rn(1.5f, ddv), dsq2) : 0.0f;
  if (!finitef(scale)) scale = 0.0f;
  const float sx = __fmul_rn(dxy_x, scale), sy = __fmul_rn(dxy_y, scale), sz = __fmul_rn(dz, scale);
  if (first) {
    const float nvx = __fadd_rn(a.vx, __fmul_rn(sx, mw1)), nvy = __fadd_rn(a.vy, __fmul_rn(sy, mw1)), nvz = __fadd_rn(a.vz, __fmul_rn(sz, mw1));
    out.dvx = __fsub_rn(nvx, a.vx), out.dvy = __fsub_rn(nvy, a.vy), out.dvz = __fsub_rn(nvz, a.vz);
    out.dx = __fsub_rn(__fadd_rn(p1x, __fmul_rn(nvx, t)), a.x);
    out.dy = __fsub_rn(__fadd_rn(p1y, __fmul_rn(nvy, t)), a.y);
    out.dz = __fsub_rn(__fadd_rn(z1, __fmul_rn(nvz, t)), a.z);
  } else {
    const float nvx = __fsub_rn(b.vx, __fmul_rn(sx, mw2)), nvy = __fsub_rn(b.vy, __fmul_rn(sy, mw2)), nvz = __fsub_rn(b.vz, __fmul_rn(sz, mw2));
    out.dvx = __fsub_rn(nvx, b.vx), out.dvy = __fsub_rn(nvy, b.vy), out.dvz = __fsub_rn(nvz, b.vz);
    out.dx = __fsub_rn(__fadd_rn(p2x, __fmul_rn(nvx, t)), b.x);
    out.dy = __fsub_rn(__fadd_rn(p2y, __fmul_rn(nvy, t)), b.y);
    out.dz = __fsub_rn(__fadd_rn(z2, __fmul_rn(nvz, t)), b.z);
  }
  return true;
}

__device__ __forceinline__ float sane(float v) { return finitef(v) ? v : 0.0f; }

// one body per thread: the sum of what its overlapping pairs ask of it
__global__ void __launch_bounds__(128)
    collide_kernel(uint32_t first, uint32_t n, const uint32_t* __restrict__ cell_start,
                   const uint32_t* __restrict__ cell_end, const uint32_t* __restrict__ body_cell,
                   const float4* __restrict__ cpos, const float4* __restrict__ recA, const float4* __restrict__ recB,
                   const uint8_t* __restrict__ species, const float4* __restrict__ accm, CollideParams P,
                   float4* __restrict__ pqr, float4* __restrict__ velz, unsigned long long* __restrict__ pair_count) {
  const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pqr[i], v = velz[i];
  // A body at a non-finite position has no bounding square to intersect: it takes part in no pair (the reference's
  // BVH is undefined for it; its resolve() would zero the coordinates of such a body if it ever met one).
  if (!(finitef(p.x) && finitef(p.y) && finitef(v.z))) return;
  CollideBody me;
  me.x = p.x, me.y = p.y, me.z = v.z, me.r = p.w, me.vx = sane(v.x), me.vy = sane(v.y), me.vz = sane(v.w);
  me.m = accm[i].w;
  me.species = species[i], me.index = i;
  const uint32_t c = body_cell[i];
  const int cx = (int)(c % P.g.gx), cy = (int)(c / P.g.gx);
  const int y0 = max(cy - 1, 0), y1 = min(cy + 1, (int)P.g.gy - 1);
  const int x0 = max(cx - 1, 0), x1 = min(cx + 1, (int)P.g.gx - 1);
  float dx = 0.f, dy = 0.f, dz = 0.f, dvx = 0.f, dvy = 0.f, dvz = 0.f;
  bool placed = false;
  float px = 0.f, py = 0.f, pz = 0.f;
  uint32_t pairs = 0;
  for (int y = y0; y <= y1; ++y) {
    for (int x = x0; x <= x1; ++x) {
      const uint32_t cc = (uint32_t)x + (uint32_t)y * P.g.gx;
      const uint32_t k1 = cell_end[cc];
      for (uint32_t k = cell_start[cc]; k < k1; ++k) {
        const float4 a4 = __ldg(&recA[k]);
        // broccoli's broad phase: the bounding squares [pos - r, pos + r] must intersect (collision.rs:113-123)
        const float rr = __fadd_rn(me.r, a4.w);
        if (!(fabsf(__fsub_rn(a4.x, me.x)) <= rr && fabsf(__fsub_rn(a4.y, me.y)) <= rr) || !finitef(a4.z)) continue;
        const float4 c4 = __ldg(&cpos[k]);
        const uint32_t j = __float_as_uint(c4.w);
        if (j == i) continue;
        const float4 b4 = __ldg(&recB[k]);
        CollideBody o;
        o.x = a4.x, o.y = a4.y, o.z = a4.z, o.r = a4.w;
        o.vx = sane(b4.x), o.vy = sane(b4.y), o.vz = sane(b4.z), o.m = b4.w;
        o.species = __float_as_uint(c4.z), o.index = j;
        CollideDelta d;
        const bool hit = i < j ? collide_resolve(me, o, true, P, d) : collide_resolve(o, me, false, P, d);
        if (!hit) continue;
        ++pairs;
        if (d.set_pos) {
          if (!placed) placed = true, px = d.px, py = d.py, pz = d.pz;  // first degenerate partner wins
        } else {
          dx = __fadd_rn(dx, d.dx), dy = __fadd_rn(dy, d.dy), dz = __fadd_rn(dz, d.dz);
          dvx = __fadd_rn(dvx, d.dvx), dvy = __fadd_rn(dvy, d.dvy), dvz = __fadd_rn(dvz, d.dvz);
        }
      }
    }
  }
  if (pairs == 0) return;
  float nx = placed ? px : me.x, ny = placed ? py : me.y, nz = placed ? pz : me.z;
  nx = __fadd_rn(nx, dx), ny = __fadd_rn(ny, dy), nz = __fadd_rn(nz, dz);
  pqr[i] = make_float4(sane(nx), sane(ny), p.z, p.w);
  velz[i] = make_float4(sane(__fadd_rn(me.vx, dvx)), sane(__fadd_rn(me.vy, dvy)), sane(nz), sane(__fadd_rn(me.vz, dvz)));
  if (pair_count) atomicAdd(pair_count, (unsigned long long)pairs);  // every touching pair is counted by both partners
}

}  // namespace psim
