// traverse.cuh — Barnes-Hut Coulomb field / force traversal.
//
// Replaces Quadtree::acc_pos (src/quadtree/quadtree.rs:350-407) driven by Quadtree::field
// (:418-427), Quadtree::field_at_point (:504-507), the `attract` epilogue
// (src/simulation/forces.rs:37-43) and the per-electron drift of Body::update_electrons
// (src/body/electron.rs:19-46).
//
// One warp walks the compact pre-order tree for 32 spatially adjacent targets (targets are in
// Morton order).  The walk visits the union of the 32 reference traversals in DFS order; every
// lane applies the reference's opening test with ITS OWN position and radius, so each target sums
// exactly the reference's interaction set, in the reference's order:
//   - a lane that accepts a node (or direct-sums a leaf) records the node's skip pointer; it ignores
//     every node below that index, i.e. the subtree it has already accounted for;
//   - the warp descends (n+1) if any still-active lane rejects an internal node, else follows the
//     skip pointer.
// Node records are two 16-byte loads at a warp-uniform address (one L1 broadcast each); descending
// is a sequential walk through memory because of the pre-order layout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tree.cuh"

namespace psim {

struct FieldParams {
  float t_sq, e_sq, k_e;
  float bg_x, bg_y;
};

// PARITY = true: IEEE sqrt/div, no FMA contraction, the reference's operation order.
// PARITY = false: rsqrt / fast divide / FMA (reported separately; still checked against the oracle).
template <bool PARITY>
struct Arith {
  static __device__ __forceinline__ float add(float a, float b) { return PARITY ? __fadd_rn(a, b) : a + b; }
  static __device__ __forceinline__ float sub(float a, float b) { return PARITY ? __fsub_rn(a, b) : a - b; }
  static __device__ __forceinline__ float mul(float a, float b) { return PARITY ? __fmul_rn(a, b) : a * b; }
  static __device__ __forceinline__ float div(float a, float b) { return PARITY ? __fdiv_rn(a, b) : __fdividef(a, b); }
  static __device__ __forceinline__ float sqrt(float a) { return PARITY ? __fsqrt_rn(a) : __fsqrt_rn(a); }
};

struct TraverseCounters {
  unsigned long long steps;  // warp steps (nodes visited by a warp)
};

// Shared walk.  (px, py, q, radius) are per-lane; `live` is false for tail lanes.
// COUNT adds per-lane interaction counters (cnt[0] = internal nodes this lane opened, cnt[1] =
// monopoles accepted, cnt[2] = direct body terms): the reference visits 1 + 4 * opened nodes.
template <bool PARITY, bool COUNT = false>
__device__ __forceinline__ float2 bh_walk(const float4* __restrict__ nodeA,
                                          const uint4* __restrict__ nodeB,
                                          const float4* __restrict__ pqr, uint32_t M, float px,
                                          float py, float q, float radius, bool live,
                                          const FieldParams P, uint32_t& steps_out,
                                          uint32_t* cnt = nullptr) {
  using A = Arith<PARITY>;
  float ax = 0.0f, ay = 0.0f;
  uint32_t skip = live ? 0u : 0xffffffffu;
  const float kq = A::mul(P.k_e, q);
  uint32_t n = 0, steps = 0;
  while (n < M) {
    const float4 na = __ldg(&nodeA[n]);
    const uint4 nb = __ldg(&nodeB[n]);
    ++steps;
    const bool active = n >= skip;
    const float dx = A::sub(px, na.x), dy = A::sub(py, na.y);
    const float d_sq = A::add(A::mul(dx, dx), A::mul(dy, dy));
    const float dist = A::sqrt(d_sq);
    const float dist_adj = fmaxf(A::sub(dist, radius), 0.0f);
    const bool accept = A::mul(na.w, na.w) < A::mul(A::mul(dist_adj, dist_adj), P.t_sq);
    const bool leaf = (nb.w & kNodeLeaf) != 0;
    if (active) {
      if (accept) {
        const float r_eff = fmaxf(dist, A::add(radius, A::mul(na.w, 0.5f)));
        const float denom = A::mul(A::add(A::mul(r_eff, r_eff), P.e_sq), r_eff);
        const float s = A::div(A::mul(kq, na.z), denom);
        ax = A::add(ax, A::mul(dx, s));
        ay = A::add(ay, A::mul(dy, s));
        skip = nb.x;
        if (COUNT) cnt[1]++;
      } else if (leaf) {
        for (uint32_t b = nb.y; b < nb.y + nb.z; ++b) {
          const float4 s4 = __ldg(&pqr[b]);
          const float ex = A::sub(s4.x, px), ey = A::sub(s4.y, py);
          if (A::add(A::mul(ex, ex), A::mul(ey, ey)) < 1e-6f) continue;  // positional self skip
          const float bx = A::sub(px, s4.x), by = A::sub(py, s4.y);
          const float bd = A::sqrt(A::add(A::mul(bx, bx), A::mul(by, by)));
          const float r_eff = fmaxf(bd, A::add(radius, s4.w));
          const float denom = A::mul(A::add(A::mul(r_eff, r_eff), P.e_sq), r_eff);
          const float s = fminf(A::div(A::mul(kq, s4.z), denom), 3.402823466e+38f);
          ax = A::add(ax, A::mul(bx, s));
          ay = A::add(ay, A::mul(by, s));
          if (COUNT) cnt[2]++;
        }
        skip = nb.x;
      } else if (COUNT) {
        cnt[0]++;
      }
    }
    const bool descend = __any_sync(0xffffffffu, active && !accept && !leaf);
    n = descend ? n + 1 : nb.x;
  }
  steps_out = steps;
  return make_float2(ax, ay);
}

// Quadtree::field + attract: targets are the bodies themselves (q_test = 1, radius_i).
template <bool PARITY>
__global__ void __launch_bounds__(128)
    bh_field_bodies_kernel(const TreeMeta* __restrict__ meta, const float4* __restrict__ nodeA,
                           const uint4* __restrict__ nodeB, const float4* __restrict__ pqr,
                           const float4* __restrict__ acc_mass /* {ax, ay, az, mass} */, uint32_t n,
                           FieldParams P, float2* __restrict__ e_field,
                           float4* __restrict__ acc_mass_out, int write_acc,
                           unsigned long long* __restrict__ step_counter) {
  const uint32_t M = meta->num_nodes;
  const uint32_t warps_per_block = blockDim.x >> 5;
  const uint32_t warp_global = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const uint32_t total_warps = gridDim.x * warps_per_block;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t groups = (n + 31) / 32;
  unsigned long long my_steps = 0;
  for (uint32_t g = warp_global; g < groups; g += total_warps) {
    const uint32_t i = g * 32 + lane;
    const bool live = i < n;
    float4 me = make_float4(0, 0, 0, 0);
    if (live) me = pqr[i];
    uint32_t steps;
    float2 e = bh_walk<PARITY>(nodeA, nodeB, pqr, M, me.x, me.y, 1.0f, me.w, live, P, steps);
    my_steps += steps;
    if (live) {
      // forces.rs:37-43: e_field += background; acc = (charge * e_field) / mass
      e.x = __fadd_rn(e.x, P.bg_x);
      e.y = __fadd_rn(e.y, P.bg_y);
      e_field[i] = e;
      if (write_acc) {
        float4 am = acc_mass[i];
        am.x = __fdiv_rn(__fmul_rn(me.z, e.x), am.w);
        am.y = __fdiv_rn(__fmul_rn(me.z, e.y), am.w);
        acc_mass_out[i] = am;
      }
    }
  }
  if (step_counter && lane == 0 && my_steps) atomicAdd(step_counter, my_steps);
}

// diagnostic: the interaction counters of Quadtree::field for the roofline's algorithmic flops
__global__ void __launch_bounds__(128)
    bh_count_bodies_kernel(const TreeMeta* __restrict__ meta, const float4* __restrict__ nodeA,
                           const uint4* __restrict__ nodeB, const float4* __restrict__ pqr, uint32_t n,
                           FieldParams P, unsigned long long* __restrict__ out4) {
  const uint32_t M = meta->num_nodes;
  const uint32_t warps_per_block = blockDim.x >> 5;
  const uint32_t warp_global = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const uint32_t total_warps = gridDim.x * warps_per_block;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t groups = (n + 31) / 32;
  unsigned long long opened = 0, acc = 0, pairs = 0, wsteps = 0;
  for (uint32_t g = warp_global; g < groups; g += total_warps) {
    const uint32_t i = g * 32 + lane;
    const bool live = i < n;
    float4 me = make_float4(0, 0, 0, 0);
    if (live) me = pqr[i];
    uint32_t steps, cnt[3] = {0, 0, 0};
    bh_walk<true, true>(nodeA, nodeB, pqr, M, me.x, me.y, 1.0f, me.w, live, P, steps, cnt);
    opened += cnt[0], acc += cnt[1], pairs += cnt[2];
    if (lane == 0) wsteps += steps;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    opened += __shfl_xor_sync(0xffffffffu, opened, off);
    acc += __shfl_xor_sync(0xffffffffu, acc, off);
    pairs += __shfl_xor_sync(0xffffffffu, pairs, off);
  }
  if (lane == 0) {
    atomicAdd(&out4[0], opened);
    atomicAdd(&out4[1], acc);
    atomicAdd(&out4[2], pairs);
    atomicAdd(&out4[3], wsteps);
  }
}

// FP32 FMA pipe peak: 8 independent FMA chains per thread (no memory traffic)
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fmaf(x0, a, b), x1 = fmaf(x1, a, b), x2 = fmaf(x2, a, b), x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b), x5 = fmaf(x5, a, b), x6 = fmaf(x6, a, b), x7 = fmaf(x7, a, b);
    }
  }
  const float r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (r == 123.456f) out[0] = r;
}

// acc_pos at arbitrary points.  q / radius may be null (1 and 0: field_at_point).
template <bool PARITY>
__global__ void __launch_bounds__(128)
    bh_field_points_kernel(const TreeMeta* __restrict__ meta, const float4* __restrict__ nodeA,
                           const uint4* __restrict__ nodeB, const float4* __restrict__ pqr,
                           const float2* __restrict__ pts, const float* __restrict__ q,
                           const float* __restrict__ radius, uint32_t m, FieldParams P,
                           float2* __restrict__ out, unsigned long long* __restrict__ step_counter) {
  const uint32_t M = meta->num_nodes;
  const uint32_t warps_per_block = blockDim.x >> 5;
  const uint32_t warp_global = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const uint32_t total_warps = gridDim.x * warps_per_block;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t groups = (m + 31) / 32;
  unsigned long long my_steps = 0;
  for (uint32_t g = warp_global; g < groups; g += total_warps) {
    const uint32_t i = g * 32 + lane;
    const bool live = i < m;
    float2 p = make_float2(0, 0);
    float qq = 1.0f, rr = 0.0f;
    if (live) {
      p = pts[i];
      if (q) qq = q[i];
      if (radius) rr = radius[i];
    }
    uint32_t steps;
    const float2 e = bh_walk<PARITY>(nodeA, nodeB, pqr, M, p.x, p.y, qq, rr, live, P, steps);
    my_steps += steps;
    if (live) out[i] = e;
  }
  if (step_counter && lane == 0 && my_steps) atomicAdd(step_counter, my_steps);
}

// ------------------------------------------------------------------------------------------------
// electrons (body/electron.rs:19-46).  Electrons are stored grouped by body, in body order, so the
// sample points of consecutive electrons are spatially adjacent.
__global__ void __launch_bounds__(256)
    electron_points_kernel(const float4* __restrict__ pqr, const uint32_t* __restrict__ ebody,
                           const float2* __restrict__ erel, uint32_t m, float2* __restrict__ pts) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += stride) {
    const float4 b = pqr[ebody[k]];
    const float2 r = erel[k];
    pts[k] = make_float2(__fadd_rn(b.x, r.x), __fadd_rn(b.y, r.y));
  }
}

struct SpeciesRow {  // mirrors psim_species in include/psim_b200.h
  float mass, radius, damping;
  float lj_epsilon, lj_sigma, lj_cutoff;
  float polar_offset, polar_charge;
  float repulsion_strength, repulsion_cutoff;
  uint32_t lj_enabled, repulsion_enabled;
};
constexpr int kMaxSpecies = 32;

__global__ void __launch_bounds__(256)
    electron_drift_kernel(const float4* __restrict__ pqr, const uint8_t* __restrict__ species,
                          const SpeciesRow* __restrict__ table, const uint32_t* __restrict__ ebody,
                          float2* __restrict__ erel, float2* __restrict__ evel,
                          const float2* __restrict__ field, uint32_t m, float bg_x, float bg_y,
                          float dt, float spring_k, float max_speed_factor) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += stride) {
    const uint32_t b = ebody[k];
    const float radius = pqr[b].w;
    const float2 f = field[k];
    const float lx = __fadd_rn(f.x, bg_x), ly = __fadd_rn(f.y, bg_y);
    const float accx = __fmul_rn(-lx, spring_k), accy = __fmul_rn(-ly, spring_k);
    float2 v = evel[k];
    v.x = __fadd_rn(v.x, __fmul_rn(accx, dt));
    v.y = __fadd_rn(v.y, __fmul_rn(accy, dt));
    const float speed = __fsqrt_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)));
    const float max_speed = __fdiv_rn(__fmul_rn(max_speed_factor, radius), dt);
    if (speed > max_speed) {
      v.x = __fmul_rn(__fdiv_rn(v.x, speed), max_speed);
      v.y = __fmul_rn(__fdiv_rn(v.y, speed), max_speed);
    }
    float2 r = erel[k];
    r.x = __fadd_rn(r.x, __fmul_rn(v.x, dt));
    r.y = __fadd_rn(r.y, __fmul_rn(v.y, dt));
    const float max_dist = __fmul_rn(table[species[b] < kMaxSpecies ? species[b] : 0].polar_offset, radius);
    const float rm = __fsqrt_rn(__fadd_rn(__fmul_rn(r.x, r.x), __fmul_rn(r.y, r.y)));
    if (rm > max_dist) {
      const float inv = __fdiv_rn(1.0f, rm);  // Vec2::normalized: multiply by 1/mag
      r.x = __fmul_rn(__fmul_rn(r.x, inv), max_dist);
      r.y = __fmul_rn(__fmul_rn(r.y, inv), max_dist);
    }
    evel[k] = v;
    erel[k] = r;
  }
}

}  // namespace psim
